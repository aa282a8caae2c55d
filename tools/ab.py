"""Build experimental variants of libicnv.so (compile-time -D switches) for A/B timing on ONE box:
    python tools/ab.py build  NAME:DEF1,DEF2 ...     (here, CPU)
    python tools/ab.py run N                          (on the GPU box; runs quick_bench per variant)"""
import sys, os, subprocess
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
VAR = ROOT / "infercnvpy_b200" / "ab"
if sys.argv[1] == "build":
    from infercnvpy_b200 import _build
    VAR.mkdir(exist_ok=True)
    for spec in sys.argv[2:]:
        name, _, defs = spec.partition(":")
        _build.build(force=True, verbose=True, defines=tuple(d for d in defs.split(",") if d), out=VAR / f"libicnv_{name}.so")
else:
    n = sys.argv[2] if len(sys.argv) > 2 else "50000"
    libs = [("product", None)] + [(p.stem.replace("libicnv_", ""), p) for p in sorted(VAR.glob("libicnv_*.so"))]
    for rep in range(2):
        for name, path in libs:
            env = dict(os.environ, QB_WINDOWS=os.environ.get("QB_WINDOWS", "100"), QB_REPS="10")
            if path: env["ICNV_LIB_PATH"] = str(path)
            r = subprocess.run([sys.executable, str(ROOT / "tools" / "quick_bench.py"), n], env=env, capture_output=True, text=True)
            for ln in r.stdout.splitlines():
                if '"window"' in ln:
                    import json; d = json.loads(ln)
                    print(f"{name:12s} w{d['window']} smooth_ms={d['smooth_ms'][0]:.4f} frac={d['smooth_frac']:.4f} center_ms={d['center_ms'][0]:.4f} colsum_ms={d['colsum_ms'][0]:.4f} thr_ms={d['thr_ms'][0]:.4f}")
            if r.returncode: print(name, "FAILED", r.stderr[-500:])
