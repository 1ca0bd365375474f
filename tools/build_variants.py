#!/usr/bin/env python
"""Build tuning / experiment variants of libparasail_b200.so into variants/ (git-ignored, but it travels
to the GPU box).  Time them against the default build with tools/variant_bench.sh, e.g.

    python tools/build_variants.py shifted
    gpurun -- 'tools/variant_bench.sh parasail_rs_b200/libparasail_b200.so variants/lib_shifted.so'

variants:
  prof8     scan kernel with the int8 profile joined by PRMT (SW16_PROF32=0)
  nopp      scan kernel without the ping-pong column copies (SW16_PINGPONG=0)
  shifted   scan kernel with round 1's shifted recurrence (three dependent instructions per row)
  wavefl    wavefront kernel, local mode: all-lanes-active step + late end cell
  wavefk    wavefront kernel, local mode: all-lanes-active step + keyed end cell
  wavekey   wavefront kernel, local mode: keyed (branch-free) end cell in score-only and traced launches
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from parasail_rs_b200 import build  # noqa: E402

VARIANTS = {
    "prof8": ["SW16_PROF32=0"],
    "nopp": ["SW16_PINGPONG=0"],
    "shifted": ["SW16_DECOUPLE=0"],
    "wavekey": ["WAVE32_SW_KEYED=3"],
    "wavefl": ["WAVE32_SW_KEYED=0", "WAVE32_SW_FSTEP=1"],
    "wavefk": ["WAVE32_SW_KEYED=3", "WAVE32_SW_FSTEP=1"],
}


def main():
    names = sys.argv[1:] or list(VARIANTS)
    os.makedirs(os.path.join(ROOT, "variants"), exist_ok=True)
    for name in names:
        out = os.path.join(ROOT, "variants", f"lib_{name}.so")
        build.build_library(force=True, defines=VARIANTS[name], out=out)
        r = subprocess.run(f"cuobjdump --dump-resource-usage {out} | grep -A1 'scan_kernelILi25' | grep REG", shell=True,
                           capture_output=True, text=True)
        print(name, out, r.stdout.strip())


if __name__ == "__main__":
    main()
