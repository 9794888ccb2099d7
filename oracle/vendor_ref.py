"""Stage the UNMODIFIED reference files of the bridge path under ``oracle/_ref/`` so that the reference arm
(``bench.py --impl reference`` / ``cpu_baseline``) runs the reference's OWN ``psd`` /
``EncoderProjectorLinearSiLU`` / ``_merge_input_ids_with_audio_features`` on the GPU box, where
``/root/reference`` does not exist.

TEST / BASELINE INFRASTRUCTURE ONLY.  ``oracle/_ref/`` is git-ignored (the reference's sources never enter this
repository's history) but travels to the GPU box with the working tree, like the built ``.so``.  Run by
``__graft_entry__.build()`` whenever ``/root/reference`` is present; a no-op elsewhere.

The files are byte-for-byte copies (their sha256 is recorded in ``oracle/_ref/MANIFEST.json``):
  model/ps-slm.py, model/projector.py        the path itself
  utils/{metric,config_utils,model_utils,dataset_utils}.py   imported at module scope by model/ps-slm.py:17-19
"""
import hashlib
import json
import os
import shutil

SRC = os.environ.get("TASU_REFERENCE_ROOT", "/root/reference/Multitask")
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "Multitask")
FILES = ["model/ps-slm.py", "model/projector.py", "utils/metric.py", "utils/config_utils.py", "utils/model_utils.py",
         "utils/dataset_utils.py"]


def vendor(verbose: bool = False) -> bool:
    """Copy the files; returns True when the staged tree is complete (already staged or just copied)."""
    if not os.path.isfile(os.path.join(SRC, FILES[0])):
        return os.path.isfile(os.path.join(DST, FILES[0]))
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(SRC, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        with open(dst, "rb") as f:
            manifest[rel] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(os.path.dirname(DST), "MANIFEST.json"), "w") as f:
        json.dump({"source": SRC, "sha256": manifest}, f, indent=1)
    if verbose:
        print("staged %d reference files under %s" % (len(FILES), DST))
    return True


if __name__ == "__main__":
    vendor(verbose=True)
