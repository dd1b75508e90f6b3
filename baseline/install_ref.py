#!/usr/bin/env python
"""Install the UNMODIFIED reference (visinf/da-sac) for the reference arm of bench.py.

    python baseline/install_ref.py            # build container only: reads /root/reference, writes baseline/_ref/

The reference is a script tree without packaging (no setup.py / pyproject.toml), so ``pip install --target baseline/_ref
/root/reference`` has nothing to build (recorded in DESIGN.md).  What its target step needs imports cleanly from a plain
copy: ``models/`` (SAC, the three backbones), ``core/config.py`` (+ ``utils/collections.py``) and ``configs/*.yaml``
(SURVEY.md 8c).  This script copies exactly those files, byte for byte, into the git-ignored ``baseline/_ref/`` and writes
``MANIFEST.json`` with their sha256 so that the GPU box (which has no /root/reference) can verify that it times the files as
they lie in the reference.  Nothing under ``baseline/_ref`` is committed; nothing in the product imports it --
``bench.py --impl reference`` and the ``cpu_baseline`` leg are its only users.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
DST = os.path.join(HERE, "_ref")
FILES = ["models/__init__.py", "models/basenet.py", "models/deeplabv2.py", "models/fcn.py", "models/sac.py",
         "core/__init__.py", "core/config.py", "utils/collections.py",
         "configs/deeplabv2_resnet101_train.yaml", "configs/deeplabv2_vgg16_train.yaml", "configs/fcn_vgg16_train.yaml",
         "LICENSE"]


def sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def install(verbose=True):
    if not os.path.isdir(REF):
        if os.path.isfile(os.path.join(DST, "MANIFEST.json")):
            return verify()
        raise SystemExit("baseline/install_ref.py: %s is not present and baseline/_ref has not been installed" % REF)
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(REF, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = sha(dst)
    json.dump({"source": "visinf/da-sac (/root/reference), unmodified", "files": manifest},
              open(os.path.join(DST, "MANIFEST.json"), "w"), indent=1, sort_keys=True)
    if verbose:
        print("installed %d reference files into %s" % (len(FILES), DST))
    return manifest


def verify():
    """sha256 of every installed file against the manifest written at install time"""
    m = json.load(open(os.path.join(DST, "MANIFEST.json")))["files"]
    bad = [rel for rel, h in m.items() if not os.path.isfile(os.path.join(DST, rel)) or sha(os.path.join(DST, rel)) != h]
    if bad:
        raise SystemExit("baseline/_ref differs from the installed reference: %s" % bad)
    return m


if __name__ == "__main__":
    install()
    sys.exit(0)
