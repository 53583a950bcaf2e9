"""A/B of build variants of the fused kernels (development aid; round-2 results: profiles/r02_ab_rejected.txt).

    python scripts/ab_variants.py build      # here, on the CPU: nvcc cross-compiles every variant into lib/variants/
    python scripts/ab_variants.py run        # on the B200 (gpurun): correctness spot check + fwd / bwd / step times per variant
    python scripts/ab_variants.py test NAME  # on the B200: the fused parity tests against one variant

The built libraries are git-ignored but travel to the GPU box with the working tree."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VARDIR = os.path.join(ROOT, "maxsquareloss_b200", "lib", "variants")
VARIANTS = {
    "base": (),
    "trace": ("MSQ_TRACE=1",),
    "tw256": ("MSQ_TW=256", "MSQ_FWD_MINB=2", "MSQ_BWD_MINB=2", "MSQ_MULTI_MINB=1", "MSQ_SRC_MINB=2"),       # measured: 28.93 vs 28.55 us
}


def path(name):
    return os.path.join(VARDIR, f"libmsq_{name}.so")


if __name__ == "__main__":
    cmd = sys.argv[1] if len(sys.argv) > 1 else "run"
    if cmd == "build":
        from maxsquareloss_b200 import build
        os.makedirs(VARDIR, exist_ok=True)
        for name, defs in VARIANTS.items():
            print(name, build.build(force=True, defines=defs + ("MSQ_VARIANT=1",), out=path(name)), flush=True)
    elif cmd == "run":
        for name in VARIANTS:
            if not os.path.exists(path(name)):
                print(f"{name}: not built"); continue
            env = dict(os.environ, MSQ_B200_LIB=path(name))
            subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ab_fused.py")], env=env)
    elif cmd == "test":
        env = dict(os.environ, MSQ_B200_LIB=path(sys.argv[2]))
        sys.exit(subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests"), "-m", "gpu", "-x", "-q", "-k",
                                 "fused or multi or source or entropy or step"], env=env).returncode)
