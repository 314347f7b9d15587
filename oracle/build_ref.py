"""Recipe for oracle/_ref: a verbatim, git-ignored copy of the reference modules the CPU arm needs (TEST / BASELINE
INFRASTRUCTURE, never on the product path).

/root/reference exists only in the build container; the GPU box receives the working tree (git-ignored files included,
see .gitignore / .gpurunignore), so `__graft_entry__.build()` runs this recipe here and `oracle/_ref/` travels with the
snapshot.  Nothing is edited: the files are byte-for-byte copies, laid out as `oracle/_ref/code/REC/...`, and
`oracle/ref_harness.py` imports them exactly as it imports /root/reference/code (same stubs for the four absent
logging-only packages).  With it `bench.py --impl reference` and `cpu_baseline` run the UNMODIFIED reference model
(`kind: "reference"`) on the GPU box's host cores instead of the restated oracle (`kind: "port"`).

Reference sources are never committed: oracle/_ref/ is listed in .gitignore.
"""
import os
import shutil

SRC = os.environ.get("B200REC_REFERENCE", "/root/reference/code")
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "code")

# import closure of REC.model.IDNet.hstu and REC.evaluator
FILES = ["REC/__init__.py", "REC/model/IDNet/hstu.py", "REC/model/basemodel.py", "REC/model/llm_heads.py",
         "REC/model/layers.py"]
DIRS = ["REC/evaluator", "REC/utils"]


def build(verbose=False):
    """Copies the files; returns True if oracle/_ref is usable afterwards (False when there is no source here)."""
    if not os.path.isfile(os.path.join(SRC, "REC", "model", "IDNet", "hstu.py")):
        return os.path.isfile(os.path.join(DST, "REC", "model", "IDNet", "hstu.py"))
    for rel in FILES:
        src = os.path.join(SRC, rel)
        if not os.path.isfile(src):
            continue
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
    for rel in DIRS:
        for name in sorted(os.listdir(os.path.join(SRC, rel))):
            if name.endswith(".py"):
                dst = os.path.join(DST, rel, name)
                os.makedirs(os.path.dirname(dst), exist_ok=True)
                shutil.copyfile(os.path.join(SRC, rel, name), dst)
    if verbose:
        n = sum(len(f) for _, _, f in os.walk(DST))
        print(f"oracle/_ref: {n} files copied from {SRC}")
    return True


if __name__ == "__main__":
    build(verbose=True)
