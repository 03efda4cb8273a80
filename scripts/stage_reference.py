"""Stage the UNMODIFIED reference for the GPU box: copy the Python files its SAM2-UNet path imports from
/root/reference into the git-ignored baseline/_ref/ (which travels with the gpurun snapshot), so
that `bench.py --impl reference` and the `cpu_baseline` leg time the reference's own modules there instead of the
oracle port.  Nothing under baseline/_ref is tracked; bench.py falls back to the port when the directory is absent."""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("SAM2UNET_REFERENCE_SRC", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")


def stage() -> bool:
    if not os.path.isfile(os.path.join(SRC, "SAM2UNet.py")):
        return False
    os.makedirs(DST, exist_ok=True)
    for name in ("SAM2UNet.py", "train.py", "dataset.py", "eval.py", "LICENSE"):
        shutil.copy2(os.path.join(SRC, name), os.path.join(DST, name))
    for sub in ("sam2", "sam2_configs"):
        dst = os.path.join(DST, sub)
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        shutil.copytree(os.path.join(SRC, sub), dst,
                        ignore=shutil.ignore_patterns("*.pyd", "*.so", "*.cu", "__pycache__", "*.pyc"))
    return True


if __name__ == "__main__":
    print("staged" if stage() else "reference not present", DST)
    sys.exit(0)
