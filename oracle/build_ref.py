"""
ORACLE TOOLING -- TEST INFRASTRUCTURE ONLY.  Recipe for `oracle/_ref/`: the UNMODIFIED reference scripts of the hot path, for

  * the reference arm of bench.py (`--impl reference`, `cpu_baseline.kind = "reference"`): the reference's own
    `train_soft_intro_vae()` driven on the host cores of the GPU box, where /root/reference does not exist;
  * tests/test_gpu_boundary.py: the reference's own `main.py` run against the drop-in module, and a checkpoint written by the
    drop-in loaded back into the reference's `SoftIntroVAE`.

The reference is a directory of Python scripts (no setup.py / pyproject, nothing to compile), so "building" it means taking
the files where they lie under /root/reference, byte for byte: `oracle/_ref/` is git-ignored (never part of the repository's
history) and is listed in MANIFEST.json with the sha256 of every file.  Nothing in the product path reads it.

Usage:  python oracle/build_ref.py      (needs /root/reference; a no-op with a message when it is absent)
"""
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("SIVAE_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(ROOT, "oracle", "_ref")
FILES = [
    "soft_intro_vae/main.py", "soft_intro_vae/train_soft_intro_vae.py", "soft_intro_vae/dataset.py",
    "soft_intro_vae_bootstrap/main.py", "soft_intro_vae_bootstrap/train_soft_intro_vae_bootstrap.py",
    "soft_intro_vae_bootstrap/dataset.py",
    # imported at the top of the trainers (`from metrics.fid_score import calculate_fid_given_dataset`, :27)
    "soft_intro_vae/metrics/fid_score.py", "soft_intro_vae/metrics/inception.py",
    "soft_intro_vae_bootstrap/metrics/fid_score.py",
    "soft_intro_vae_bootstrap/metrics/inception.py",
]


def build(verbose=True):
    if not os.path.isdir(REF):
        if verbose:
            print("oracle/build_ref.py: %s not present -- keeping oracle/_ref as it is (%s)" % (
                REF, "present" if os.path.isdir(OUT) else "absent"))
        return os.path.isdir(OUT)
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(REF, rel), os.path.join(OUT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    json.dump(dict(source=REF, files=manifest), open(os.path.join(OUT, "MANIFEST.json"), "w"), indent=1)
    if verbose:
        print("oracle/_ref: %d reference files copied unmodified from %s" % (len(manifest), REF))
    return True


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
