"""Stage the reference's Python sources for the GPU box (TEST INFRASTRUCTURE ONLY).

    python oracle/stage_ref.py            # /root/reference/{model,config} -> oracle/_ref/{model,config}

`/root/reference` exists only in the build container.  The GPU box receives a snapshot of this repo, and `oracle/_ref/`
is git-ignored (it never enters the history) but NOT gpurun-ignored, so whatever this recipe puts there travels with the
snapshot exactly like the built `.so` files do.  It lets the `-m gpu` tests and `bench.py`'s second baseline execute the
reference's OWN, UNMODIFIED files on the B200 (`oracle/ref_callers.py`):

  * `model/predictors/InstancePredictorBase.py` (`forward_articulation`, `get_bones`, constraint masks) and
    `model/models/AnimalModel.py` (`render`), `model/models/Fauna.py` (`get_random_view_mask`) running on top of
    `3danimals_b200.overlay` - the drop-in proven under its real callers;
  * `model/geometry/{dmtet,skinning}.py`, `model/render/mesh.py` as the torch-op GPU baseline of the geometry half.

Nothing is edited: files are copied byte for byte and a manifest (relative path -> sha256) is written next to them so a test
can assert that what it executed is what the reference ships.  Nothing here is imported by the product.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SRC = os.environ.get("B2A_REFERENCE_ROOT", "/root/reference")
TREES = ("model", "config")


def _sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def available():
    """True when a staged copy (or the reference tree itself) can be used."""
    return os.path.isfile(os.path.join(root(), "model", "geometry", "dmtet.py"))


def root():
    """The staged tree when present, else the reference tree (build container), else the (absent) staged path."""
    if os.path.isfile(os.path.join(DEST, "model", "geometry", "dmtet.py")):
        return DEST
    if os.path.isfile(os.path.join(SRC, "model", "geometry", "dmtet.py")):
        return SRC
    return DEST


def stage(force=False):
    if not os.path.isdir(os.path.join(SRC, "model")):
        return None                       # GPU box: nothing to stage from, the snapshot already carries oracle/_ref
    manifest = {}
    for tree in TREES:
        src_tree = os.path.join(SRC, tree)
        for dirpath, dirnames, filenames in os.walk(src_tree):
            dirnames[:] = [d for d in dirnames if d not in ("__pycache__", "c_src")]
            for fn in filenames:
                if not fn.endswith((".py", ".yaml", ".yml", ".json", ".txt")):
                    continue
                s = os.path.join(dirpath, fn)
                rel = os.path.relpath(s, SRC)
                d = os.path.join(DEST, rel)
                os.makedirs(os.path.dirname(d), exist_ok=True)
                digest = _sha(s)
                if force or not os.path.isfile(d) or _sha(d) != digest:
                    shutil.copyfile(s, d)
                manifest[rel] = digest
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=0, sort_keys=True)
    return DEST


def verify():
    """-> list of staged files whose bytes differ from the manifest (empty = the staged tree is the reference's)."""
    mpath = os.path.join(DEST, "MANIFEST.json")
    if not os.path.isfile(mpath):
        return ["MANIFEST.json missing"]
    manifest = json.load(open(mpath))
    return [rel for rel, digest in manifest.items() if not os.path.isfile(os.path.join(DEST, rel)) or _sha(os.path.join(DEST, rel)) != digest]


if __name__ == "__main__":
    out = stage(force="--force" in sys.argv)
    print("staged" if out else "no reference tree here", out or "")
