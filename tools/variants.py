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
    "stream_all": ["-DSKA_WS_STREAM_ALL"],                                   # V <= 4 through the streaming three-pass form too
    "stream_u2": ["-DSKA_STREAM_UNROLL=2"],                                  # view loops unrolled by 2 (V >= 5)
    "stream_w23": ["-DSKA_WS_STREAM_WARPS=23"],                              # 24 warps x 85 registers (V >= 5)
    "stream_w11": ["-DSKA_WS_STREAM_WARPS=11"],                              # 12 warps x 168 registers (V >= 5)
    "stream_7x2": ["-DSKA_WS_STREAM_WARPS=7", "-DSKA_WS_STREAM_MINB=2"],     # 2 CTAs x (7 + 1) warps
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
