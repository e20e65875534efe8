"""Recipe for ``oracle/_ref/``: a build-time, git-ignored copy of the three UNMODIFIED reference files of the hot path.

The reference is interpreted Python, so "building" it is copying the sources where they lie under /root/reference
into ``oracle/_ref/`` (listed in .gitignore, NOT in .gpurunignore: it travels to the GPU box like a built .so and never
enters the history).  With it the reference arm of ``bench.py`` (``--impl reference``) and ``cpu_baseline`` time the
reference's own modules (``kind: "reference"``) instead of the restatement in ``oracle/bidatenet_oracle.py``.

    python oracle/build_ref.py        (also run by __graft_entry__.build())
"""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
FILES = ("models/bidate_model.py", "models/unet_parts.py", "utils/metrics.py")


def build(reference="/root/reference", out=None, quiet=False):
    out = out or os.path.join(HERE, "_ref")
    reference = os.environ.get("FABRIC_REFERENCE", reference)
    if not all(os.path.exists(os.path.join(reference, f)) for f in FILES):
        if not quiet:
            print(f"[oracle/_ref] {reference} not present: keeping whatever oracle/_ref already holds")
        return False
    lines = []
    for f in FILES:
        dst = os.path.join(out, f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(reference, f), dst)
        lines.append(f"{hashlib.sha256(open(dst, 'rb').read()).hexdigest()}  {f}")
    with open(os.path.join(out, "SHA256SUMS"), "w") as fh:
        fh.write("\n".join(lines) + "\n")
    if not quiet:
        print(f"[oracle/_ref] copied {len(FILES)} unmodified reference files from {reference}")
    return True


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
