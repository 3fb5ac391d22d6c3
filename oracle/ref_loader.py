"""TEST INFRASTRUCTURE — import the UNMODIFIED reference from /root/reference/src.

Works only where /root/reference exists (the build container); the GPU box has
no reference tree, so `-m gpu` tests, smoke() and bench.py never call this.
"""
from __future__ import annotations

import os
import sys

REFERENCE_SRC = "/root/reference/src"
SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_SRC, "alphazero_implementation"))


def load_reference():
    """Return a namespace with the reference's hot-path classes (imported, not copied)."""
    if not reference_available():
        raise RuntimeError("reference tree not present at /root/reference (expected on the GPU box)")
    for p in (SHIMS, REFERENCE_SRC):
        if p not in sys.path:
            sys.path.insert(0, p)
    from types import SimpleNamespace

    from alphazero_implementation.core.search.mcts import AlphaZeroSearch, Node
    from alphazero_implementation.core.training.episode import Episode, Sample
    from alphazero_implementation.core.training.episode_generator import EpisodeGenerator
    from alphazero_implementation.models.base import Model
    from alphazero_implementation.models.games.connect4 import BasicNN, CNNModel
    from simulator.game.connect import Action, Config, State

    return SimpleNamespace(
        AlphaZeroSearch=AlphaZeroSearch,
        Node=Node,
        Episode=Episode,
        Sample=Sample,
        EpisodeGenerator=EpisodeGenerator,
        Model=Model,
        BasicNN=BasicNN,
        CNNModel=CNNModel,
        Action=Action,
        Config=Config,
        State=State,
    )


def load_shim_game():
    """The `simulator.game.connect` stand-in alone (travels to the GPU box)."""
    if SHIMS not in sys.path:
        sys.path.insert(0, SHIMS)
    import simulator.game.connect as connect

    return connect
