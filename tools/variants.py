"""Build experimental variants of libska.so (different launch bounds / points per thread) so one
gpurun call can time them side by side:  python tools/variants.py build ; python tools/variants.py run"""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
VARIANTS = {
    "sw9": ["-DSKA_WS_STREAM_WARPS=9"],     # V = 6, 8 view-pair ws kernel: 10 warps x <= 204 registers
    "sw10": ["-DSKA_WS_STREAM_WARPS=10"],   # 11 warps x <= 186
    "sw13": ["-DSKA_WS_STREAM_WARPS=13", "-DSKA_WS_STREAM_STAGES=2"],   # 14 warps x <= 146
    "nowsl": ["-DSKA_NO_WS_LARGE"],         # V >= 5 in tri_kernel (view pairs, register prefetch)
}
LIBDIR = ROOT / "skiing_analysis_pytorch_b200" / "lib"


def main():
    cmd = sys.argv[1]
    names = sys.argv[2:] or list(VARIANTS)
    if cmd == "build":
        from skiing_analysis_pytorch_b200 import build

        for n in names:
            print(n, build.build(out=LIBDIR / f"libska_{n}.so", extra_flags=VARIANTS[n]))
    else:
        extra = os.environ.get("BENCH_ARGS", "--steps 20 --warmup 3 --no-cpu-baseline").split()
        for n in names:
            env = dict(os.environ, SKA_LIB_PATH=str(LIBDIR / f"libska_{n}.so"))
            r = subprocess.run([sys.executable, str(ROOT / "bench.py"), *extra], env=env, capture_output=True, text=True)
            try:
                j = json.loads(r.stdout.strip().splitlines()[-1])
                print(f"{n:12s} ms/step {j['ms_per_step']:.4f}  frac {j['roofline']['frac']:.3f}")
            except Exception:
                print(n, "FAILED", r.stdout[-300:], r.stderr[-300:])


if __name__ == "__main__":
    main()
