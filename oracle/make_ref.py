"""TEST INFRASTRUCTURE ONLY.  Recipe for oracle/_ref/: the reference's own hot-path files, taken unmodified from the
reference tree where it lies (/root/reference, build container only), so that `bench.py --impl reference` and the
oracle-pinning tests can run THE REFERENCE ITSELF on a box where /root/reference does not exist.

    python oracle/make_ref.py          (also run by __graft_entry__.build() when the reference tree is present)

oracle/_ref/ is git-ignored (no reference source enters the history) but not gpurun-ignored (it travels to the GPU box
like the built .so).  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may use it -- as the checker / the
reported baseline, never on the product path."""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
# the files the vanilla hot path lives in (the mip variant needs nerfstudio: not runnable)
FILES = ["models/__init__.py", "models/rendering__.py", "models/star__.py", "models/nerf.py", "models/resnet.py",
         "models/embedder.py", "models/types__.py", "utils/__init__.py", "utils/constants.py"]


def make(src="/root/reference"):
    if not os.path.isfile(os.path.join(src, "models", "rendering__.py")):
        return None
    for f in FILES:
        s, d = os.path.join(src, f), os.path.join(DST, f)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        if os.path.isfile(s):
            shutil.copyfile(s, d)
        else:                      # an absent package marker: an empty one serves
            open(d, "w").close()
    return DST


if __name__ == "__main__":
    print(make(sys.argv[1] if len(sys.argv) > 1 else "/root/reference"))
