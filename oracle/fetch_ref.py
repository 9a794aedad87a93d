"""Recipe that makes the UNMODIFIED reference importable on the GPU box (test / baseline infrastructure, not product).

The reference (sriramelango/optimized-diffusion-model, `/root/reference`) is pure Python: there is nothing to compile.
`/root/reference` exists only in the build container, so this script copies the files of the hot path -- byte for byte,
no edits -- into the git-ignored `oracle/_ref/Reflected-Diffusion/` (listed in .gitignore, NOT in .gpurunignore, so it
travels to the GPU box with the snapshot like the built .so files do; it never enters the history).  Consumers:
  * `bench.py --impl reference`   stock `sampling.get_sampling_fn` on the box's host cores (BASELINE config C1 in full)
  * `bench.py` (our arm)          `eager_gpu_baseline`: the same stock code with device='cuda', fp32, TF32 off
  * tests/ and smoke()            may import it as a second checker next to oracle/rd_oracle.py
Nothing under optimized-diffusion-model_b200/ may import it.

Copied: sampling.py, sde_lib.py, cube.py, losses.py, utils.py, datasets.py (codec constants) and models/*.py
(`models/__init__.py` imports every model family, so the whole package is needed to import `models.ncsnpp`).
A manifest with sha256 sums is written next to the copy so a stale / edited copy is detectable.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/Reflected-Diffusion"
DST = os.path.join(HERE, "_ref", "Reflected-Diffusion")
FILES = ["sampling.py", "sde_lib.py", "cube.py", "losses.py", "utils.py", "datasets.py"]
# the reference's main CONSUMER of the hot path (tests/test_gpu_round2.py runs its unmodified generate_samples on top of
# the B200 package): copied next to the package directory, like in the reference tree
EXTRA = {"../Benchmark/gto_halo_benchmarking.py": "../Benchmark/gto_halo_benchmarking.py"}


def _sha(path: str) -> str:
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def available() -> bool:
    return os.path.exists(os.path.join(DST, "MANIFEST.json"))


def fetch(force: bool = False) -> str:
    """Copy the reference's hot-path sources into oracle/_ref (no-op when /root/reference is absent)."""
    if not os.path.isdir(SRC):
        return DST if available() else ""
    files = list(FILES) + [os.path.join("models", f) for f in sorted(os.listdir(os.path.join(SRC, "models")))
                           if f.endswith(".py")]
    manifest = {}
    for rel in files + list(EXTRA):
        s, d = os.path.normpath(os.path.join(SRC, rel)), os.path.normpath(os.path.join(DST, EXTRA.get(rel, rel)))
        os.makedirs(os.path.dirname(d), exist_ok=True)
        if force or not os.path.exists(d) or _sha(d) != _sha(s):
            shutil.copyfile(s, d)
        manifest[rel] = _sha(d)
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": SRC, "files": manifest}, f, indent=1, sort_keys=True)
    return DST


def verify() -> bool:
    """True when every copied file still has the recorded hash (i.e. the copy is the unmodified reference)."""
    if not available():
        return False
    with open(os.path.join(DST, "MANIFEST.json")) as f:
        m = json.load(f)["files"]
    return all(os.path.exists(os.path.normpath(os.path.join(DST, rel))) and _sha(os.path.normpath(os.path.join(DST, rel))) == h
               for rel, h in m.items())


class imported:
    """Context manager: `with fetch_ref.imported() as ref:` gives the stock modules (ref.sampling, ref.sde_lib, ref.cube,
    ref.mutils, ref.ncsnpp, ref.losses) imported from oracle/_ref under their own top-level names, and restores
    sys.path / sys.modules afterwards so the drop-in modules of the same names are not shadowed."""
    NAMES = ("sampling", "sde_lib", "cube", "losses", "utils", "datasets", "models")

    def __enter__(self):
        if not available():
            raise RuntimeError("oracle/_ref is missing: run `python oracle/fetch_ref.py` in the build container")
        os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
        self._saved = {k: v for k, v in sys.modules.items() if k.split(".")[0] in self.NAMES}
        for k in list(self._saved):
            del sys.modules[k]
        sys.path.insert(0, DST)
        import types
        ref = types.SimpleNamespace()
        import sampling, sde_lib, cube  # noqa: E401
        from models import utils as mutils, ncsnpp
        ref.sampling, ref.sde_lib, ref.cube, ref.mutils, ref.ncsnpp = sampling, sde_lib, cube, mutils, ncsnpp
        assert os.path.abspath(sampling.__file__).startswith(DST), sampling.__file__
        self.ref = ref
        return ref

    def __exit__(self, *exc):
        sys.path.remove(DST)
        for k in [k for k in sys.modules if k.split(".")[0] in self.NAMES]:
            del sys.modules[k]
        sys.modules.update(self._saved)
        return False


if __name__ == "__main__":
    out = fetch(force="--force" in sys.argv)
    print(out or "reference not present here and no previous copy", "verified" if verify() else "UNVERIFIED")
