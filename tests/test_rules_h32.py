"""CPU: the half-word rules the streaming rules kernels run (csrc/c4_bitboard.cuh, namespace c4::h32) against the 64-bit formulation of
the same header (what the tree kernels run, itself pinned to the goldens through oracle/c4_oracle.c and the GPU tests), compiled for the
host by g++: tests/native/rules_h32_check.cpp."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_half_word_rules_equal_the_64_bit_rules(tmp_path):
    exe = tmp_path / "rules_h32_check"
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", str(exe), os.path.join(ROOT, "tests", "native", "rules_h32_check.cpp")], check=True)
    out = subprocess.run([str(exe), "60000"], check=True, capture_output=True, text=True).stdout
    n = [int(w) for w in out.replace(",", " ").split() if w.isdigit()]
    # positions, winning moves, drawing moves, moves into a full column: every case is in the sample
    assert n[0] == 60000 * 60 and n[1] > 10000 and n[2] > 5 and n[3] > 10000, out
