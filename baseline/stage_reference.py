"""Stage the UNMODIFIED reference under baseline/_ref (git-ignored, travels to the GPU box with gpurun).

The reference has neither setup.py nor pyproject.toml, so `pip install --target baseline/_ref /root/reference`
answers "Directory is not installable" (recorded in DESIGN.md): the install step is this byte-for-byte copy of the
python files of the path (main.py, evaluation.py, model/, modules/, utils/).  Nothing under baseline/_ref is
product source; it is never committed (see .gitignore), only executed by the reference arm of bench.py, by
tools/run_reference_main.py and by the side-by-side GPU test.  A manifest with the sha256 of every file is
written next to the copy so a run on the GPU box can state exactly which reference it timed.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
SRC = "/root/reference"
KEEP = ("__init__.py", "main.py", "evaluation.py", "model", "modules", "utils", "requirements.txt")


def stage(src=SRC, dst=DST, quiet=False):
    if not os.path.isdir(src):
        if not quiet:
            print(f"[stage_reference] {src} absent: keeping {dst} as it is")
        return os.path.exists(os.path.join(dst, "main.py"))
    os.makedirs(dst, exist_ok=True)
    manifest = {}
    for name in KEEP:
        s = os.path.join(src, name)
        if os.path.isdir(s):
            for fn in sorted(os.listdir(s)):
                if fn.endswith(".py"):
                    os.makedirs(os.path.join(dst, name), exist_ok=True)
                    shutil.copyfile(os.path.join(s, fn), os.path.join(dst, name, fn))
                    manifest[f"{name}/{fn}"] = hashlib.sha256(open(os.path.join(s, fn), "rb").read()).hexdigest()
        elif os.path.exists(s):
            shutil.copyfile(s, os.path.join(dst, name))
            manifest[name] = hashlib.sha256(open(s, "rb").read()).hexdigest()
    json.dump(manifest, open(os.path.join(dst, "MANIFEST.json"), "w"), indent=1, sort_keys=True)
    if not quiet:
        print(f"[stage_reference] {len(manifest)} files -> {dst}")
    return True


def ref_root():
    """Directory of the unmodified reference: /root/reference in the build container, baseline/_ref on the GPU box."""
    if os.path.exists(os.path.join(SRC, "main.py")):
        return SRC
    if os.path.exists(os.path.join(DST, "main.py")):
        return DST
    return None


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
