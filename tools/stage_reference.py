"""Stage the UNMODIFIED reference files of the hot path into the git-ignored baseline/_ref/ (run in the build container,
where /root/reference exists; `__graft_entry__.build()` calls this).  baseline/_ref/ is not tracked and not
gpurun-ignored, so it travels to the GPU box with the snapshot and `bench.py --impl reference` can time the reference's
own `TauL.sample` / `LBJF.sample` / `calc_loss` there (cpu_baseline.kind = "reference").  Nothing under baseline/_ref/ is
imported by the product or by the tests.

  python tools/stage_reference.py [/root/reference]
"""
import filecmp
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILES = [
    "lib/sampling/sampling.py", "lib/sampling/sampling_utils.py", "lib/sampling/__init__.py",
    "lib/models/forward_model.py", "lib/models/model_utils.py", "lib/models/__init__.py",
    "lib/losses/losses.py", "lib/losses/losses_utils.py", "lib/losses/__init__.py",
    "lib/utils/utils.py", "lib/utils/__init__.py",
    "lib/d3pm.py", "lib/d3pm_utils.py", "lib/__init__.py",      # lib/losses/losses.py imports lib.d3pm at module level
]


def stage(ref_root="/root/reference"):
    src_root = os.path.join(ref_root, "TAUnSDDM")
    if not os.path.isdir(src_root):
        return None
    dst_root = os.path.join(ROOT, "baseline", "_ref", "TAUnSDDM")
    n = 0
    for rel in FILES:
        src, dst = os.path.join(src_root, rel), os.path.join(dst_root, rel)
        if not os.path.exists(src):
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not (os.path.exists(dst) and filecmp.cmp(src, dst, shallow=False)):
            shutil.copyfile(src, dst)
        n += 1
    return dst_root if n else None


if __name__ == "__main__":
    print(stage(sys.argv[1] if len(sys.argv) > 1 else "/root/reference"))
