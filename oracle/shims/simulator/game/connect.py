"""TEST INFRASTRUCTURE — stand-in for the third-party ``simulator.game.connect``.

The reference (`/root/reference/requirements.txt:9`) pins
``git+https://github.com/jojolebarjos/board-game-simulator-python.git@0.0.4``, a
C++ extension whose source is NOT under `/root/reference`, not installed and not
in the offline wheelhouse.  This module restates only the call-site contract the
reference relies on so the reference's *own* search /
self-play code can be executed unchanged as the parity oracle.

PARITY UNPINNED for this layer: no reference test or fixture pins the rules.
Assumptions (stated in DESIGN.md): `state.actions` ascending by column; reward
`[+1,-1]` / `[-1,+1]` on a win, `[0,0]` on a draw; `has_ended` iff win or full
board; grid row 0 = bottom (`notebooks/policy_comparison.ipynb#cell6` comments),
-1 empty / 0 / 1 owner.  Cross-checked in tests against the in-tree rules of
`src/alphazero_simple/connect4_game.py:28-98` (after row flip / re-encoding).

Written grid-based (numpy cell scans) on purpose: it is an independent statement
of the rules from the product's bitboard kernel, so agreement between the two is
evidence, not tautology.  Only `tests/`, `bench.py --impl reference`/cpu_baseline
and `__graft_entry__.smoke()` may import anything under `oracle/`.
"""
from __future__ import annotations

import numpy as np


class Config:
    """Connect-k configuration (`scripts/train.py:12`: ``Config(6, 7, 4)``)."""

    def __init__(self, height: int = 6, width: int = 7, count: int = 4):
        self.height = int(height)
        self.width = int(width)
        self.count = int(count)
        self.num_players = 2

    def sample_initial_state(self) -> "State":
        grid = np.full((self.height, self.width), -1, dtype=np.int8)
        return State(self, grid, 0)

    def to_json(self):
        return {"count": self.count, "height": self.height, "width": self.width}

    def __eq__(self, other):
        return isinstance(other, Config) and self.to_json() == other.to_json()

    def __hash__(self):
        return hash((self.height, self.width, self.count))


_DIRS = ((0, 1), (1, 0), (1, 1), (1, -1))


def _winner_of(grid: np.ndarray, count: int) -> int:
    """Owner (0/1) of any `count`-in-line, else -1.  Cell scan, 4 directions."""
    h, w = grid.shape
    for r in range(h):
        for c in range(w):
            who = grid[r, c]
            if who < 0:
                continue
            for dr, dc in _DIRS:
                rr, cc, n = r, c, 1
                while n < count:
                    rr += dr
                    cc += dc
                    if rr < 0 or rr >= h or cc < 0 or cc >= w or grid[rr, cc] != who:
                        break
                    n += 1
                if n >= count:
                    return int(who)
    return -1


class State:
    __slots__ = ("config", "_grid", "player", "_winner", "_full")

    def __init__(self, config: Config, grid: np.ndarray, player: int):
        self.config = config
        self._grid = grid
        self.player = int(player)
        self._winner = _winner_of(grid, config.count)
        self._full = bool((grid[config.height - 1] >= 0).all())

    # -- observation ---------------------------------------------------------
    @property
    def grid(self) -> np.ndarray:
        return self._grid

    @property
    def has_ended(self) -> bool:
        return self._winner >= 0 or self._full

    @property
    def reward(self) -> np.ndarray:
        r = np.zeros(2, dtype=np.float32)
        if self._winner >= 0:
            r[self._winner] = 1.0
            r[1 - self._winner] = -1.0
        return r

    @property
    def actions(self) -> list["Action"]:
        if self.has_ended:
            return []
        top = self.config.height - 1
        return [Action(self, c) for c in range(self.config.width) if self._grid[top, c] < 0]

    # -- (de)serialisation, shape per notebooks/episode_generation_testing.ipynb#cell2
    def to_json(self):
        return {
            "config": self.config.to_json(),
            "grid": [[int(v) for v in row] for row in self._grid],
            "player": self.player,
        }

    @classmethod
    def from_json(cls, data) -> "State":
        cfg = data["config"]
        config = Config(cfg["height"], cfg["width"], cfg["count"])
        grid = np.array(data["grid"], dtype=np.int8).reshape(config.height, config.width)
        return cls(config, grid, data["player"])

    def _key(self):
        return (self.player, self._grid.tobytes())

    def __eq__(self, other):
        return isinstance(other, State) and self._key() == other._key()

    def __hash__(self):
        return hash(self._key())

    def __repr__(self):
        return f"State(player={self.player}, grid=\n{self._grid[::-1]})"


class Action:
    __slots__ = ("state", "column")

    def __init__(self, state: State, column: int):
        self.state = state
        self.column = int(column)

    def sample_next_state(self) -> State:
        s = self.state
        g = s._grid.copy()
        col = g[:, self.column]
        empties = np.nonzero(col < 0)[0]
        if len(empties) == 0:
            raise ValueError(f"column {self.column} is full")
        g[empties[0], self.column] = s.player
        return State(s.config, g, 1 - s.player)

    def to_json(self):
        return {"column": self.column}

    @classmethod
    def from_json(cls, data, state: State | None = None) -> "Action":
        return cls(state, data["column"])  # type: ignore[arg-type]

    def __eq__(self, other):
        return (
            isinstance(other, Action)
            and self.column == other.column
            and (self.state is other.state or self.state == other.state)
        )

    def __hash__(self):
        return hash(("A", self.column, None if self.state is None else self.state._key()))

    def __repr__(self):
        return f"Action(column={self.column})"
