#!/usr/bin/env python
"""Recipe for the reference arm of bench.py (``--impl reference``): put the UNMODIFIED reference where the GPU box
can import it.  Run in the build container (``/root/reference`` is not mounted on the GPU box):

    python baseline/install_ref.py

1. the contract's install, ``pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target
   baseline/_ref /root/reference`` -- fails: the reference is a directory of scripts without setup.py/pyproject.toml;
2. therefore: import the hot path through ``oracle/refshim.py`` from ``/root/reference``, list the modules that import
   actually loaded, and copy exactly those files (byte for byte, relative layout kept) into ``baseline/_ref/``.

``baseline/_ref/`` is git-ignored (nothing of the reference enters the history) but NOT gpurun-ignored, so it travels
to the GPU box with the snapshot; ``bench.py --impl reference`` imports it through ``oracle/refshim.py`` with
``SPB200_REFERENCE=baseline/_ref`` and falls back to the oracle port (labelled ``kind: "port"``) when it is absent.
"""
import hashlib
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEST = os.path.join(ROOT, "baseline", "_ref")
SRC = "/root/reference"


def main():
    if not os.path.isfile(os.path.join(SRC, "export.py")):
        print(f"{SRC} is not mounted: nothing to do (bench.py --impl reference will time the oracle port)")
        return 1
    log = {}
    r = subprocess.run([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--find-links",
                        "/opt/wheelhouse", "--target", DEST + ".pip", SRC], capture_output=True, text=True)
    log["pip_install_rc"] = r.returncode
    log["pip_install_tail"] = (r.stderr or r.stdout).strip().splitlines()[-1:] if r.returncode else []
    shutil.rmtree(DEST + ".pip", ignore_errors=True)
    sys.path.insert(0, ROOT)
    sys.dont_write_bytecode = True
    os.environ["SPB200_REFERENCE"] = SRC
    from oracle import refshim
    refshim.load()
    files = sorted({os.path.relpath(m.__file__, SRC) for m in list(sys.modules.values())
                    if getattr(m, "__file__", None) and os.path.abspath(m.__file__).startswith(SRC + os.sep)})
    shutil.rmtree(DEST, ignore_errors=True)
    digest = {}
    for rel in files:
        dst = os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, rel), dst)
        with open(dst, "rb") as f:
            digest[rel] = hashlib.sha256(f.read()).hexdigest()[:16]
    log["files"] = digest
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as f:
        json.dump(log, f, indent=1)
    print(f"copied {len(files)} reference files into {DEST} (pip install rc={log['pip_install_rc']})")
    return 0


if __name__ == "__main__":
    sys.exit(main())
