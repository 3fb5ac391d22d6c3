"""TEST INFRASTRUCTURE — parity oracle for the self-play / MCTS hot path.

Nothing under `oracle/` is product code.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference`
legs may import it, and only as the checker or the timed CPU baseline — never as
a fallback for the CUDA path (the product raises if its CUDA library is missing).

Contents
  shims/          stand-ins for the two packages the reference imports but this
                  image lacks (`simulator` 0.0.4 — third-party C++, source absent;
                  `lightning`), so the reference's OWN `search.py` / `node.py` /
                  `episode_generator.py` run unchanged from /root/reference/src.
  ref_loader.py   puts /root/reference/src + shims on sys.path (this container only).
  evaluators.py   deterministic evaluators (uniform prior / board-hash) with the
                  reference `Model.predict` signature.
  gen_golden.py   runs the reference itself and writes tests/golden/*.json.
  c4_oracle.c     C restatement of the reference algorithm (travels to the GPU box;
                  pinned against the goldens above in tests/test_oracle_vs_golden.py).
  c4oracle.py     ctypes wrapper + builder for c4_oracle.c.

Parity status: tree search / PUCT / backup / policy targets / move sampling /
episode assembly are pinned against outputs of the reference's own code run
here (tests/golden/, generator committed).  The game-rules layer is PARITY
UNPINNED (third-party `simulator` source absent) — see shims/simulator/game/connect.py.
"""
