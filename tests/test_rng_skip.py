"""rng_skip.h: advancing the generator without computing the sample leaves it exactly where a draw of a fresh std::normal_distribution would."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = r'''
#include <cstdio>
#include <random>
#include "rng_skip.h"
int main() {
  for (unsigned seed : {1u, 42u, 20240923u}) {
    std::mt19937 a(seed), b(seed);
    for (int i = 0; i < 2000000; ++i) {
      std::normal_distribution<double> d(0.0, 0.37);   // a new object per draw, like random::normal (src/random.cpp:69-73)
      (void)d(a);
      lama_b200::rng_skip_normal(b);
      if ((i & 1023) == 0 && a != b) { std::printf("diverged at draw %d of seed %u\n", i, seed); return 1; }
    }
    if (a != b || a() != b()) { std::printf("diverged at the end of seed %u\n", seed); return 1; }
  }
  std::printf("aligned\n");
  return 0; }
'''


def test_rng_skip_keeps_the_stream_aligned(tmp_path):
    src = tmp_path / "skip.cpp"
    src.write_text(SRC)
    exe = tmp_path / "skip"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-I", os.path.join(ROOT, "iris_lama_b200", "csrc"), str(src), "-o", str(exe)])
    assert subprocess.check_output([str(exe)]).decode().strip() == "aligned"
