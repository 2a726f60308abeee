"""Diagnostic: build variants of libctdd_b200.so that differ in the defines of ONE source file, and time them.

  python tools/variants.py build TAG SOURCE.cu [DEFINE ...]     (here, no GPU: nvcc cross-compiles)
  python tools/variants.py run PROBE.py [TAG ...]               (on the GPU box: runs the probe once per variant)

Variants live in <package>/build/variants/libctdd_<TAG>.so (git-ignored, shipped by gpurun); the probe is run with
CTDD_B200_LIB pointing at the variant.  Not part of the product.
"""
import os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "continuous-time-diffusion-models-for-discrete-data_b200")
VAR = os.path.join(PKG, "build", "variants")
sys.path.insert(0, PKG)


def build(tag, source, defines):
    import build as b
    b.build()
    os.makedirs(VAR, exist_ok=True)
    obj = os.path.join(VAR, f"{tag}.o")
    src = os.path.join(b.CSRC, source)
    cmd = [b.NVCC] + b.NVCC_FLAGS + ["-D" + d for d in defines] + ["-c", src, "-o", obj]
    subprocess.run(cmd, check=True)
    objs = [os.path.join(b.BUILD, os.path.basename(s)[:-3] + ".o") for s in b._sources() if os.path.basename(s) != source]
    lib = os.path.join(VAR, f"libctdd_{tag}.so")
    subprocess.run([b.NVCC, "-shared", "-o", lib] + objs + [obj, "-lcuda"], check=True)
    os.remove(obj)
    print(lib)


def run(probe, tags):
    if not tags:
        tags = sorted(f[len("libctdd_"):-3] for f in os.listdir(VAR) if f.startswith("libctdd_") and f.endswith(".so"))
    for t in ["product"] + tags:
        env = dict(os.environ)
        if t != "product":
            env["CTDD_B200_LIB"] = os.path.join(VAR, f"libctdd_{t}.so")
        print(f"=== {t}", flush=True)
        subprocess.run(["timeout", "300", sys.executable, probe], env=env)


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build(sys.argv[2], sys.argv[3], sys.argv[4:])
    else:
        run(sys.argv[2], sys.argv[3:])
