"""Recipe for ``oracle/_ref/``: the UNMODIFIED reference renderer, vendored from /root/reference into a git-ignored
directory so that it travels to the GPU box like the built ``.so`` files (the reference is pure Python: there is
nothing to compile; "building" it is copying the two files of the hot path byte for byte).

    python -m oracle.build_ref        # also run by __graft_entry__.build() when /root/reference is present

Copies nerfServer/core/run_nerf_helpers.py and nerfServer/core/run_nerf.py (get_rays, Embedder, NeRF, sample_pdf,
render, batchify_rays, render_rays, raw2outputs, run_network: rows a2-a10 of SURVEY.md section 8) and writes a
manifest with their SHA-256.  ``oracle/ref_import.py`` imports them from /root/reference when that exists, else from
here; ``bench.py --impl reference`` then times the reference itself (``cpu_baseline.kind = "reference"``) instead
of the oracle port.  Nothing under oracle/_ref/ is ever committed (.gitignore) or imported by the product.

TEST / BENCH INFRASTRUCTURE.
"""
import hashlib
import json
import os
import shutil

SRC = "/root/reference/nerfServer/core"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "nerfServer", "core")
FILES = ["run_nerf_helpers.py", "run_nerf.py"]


def build():
    if not os.path.isdir(SRC):
        return None
    os.makedirs(DST, exist_ok=True)
    manifest = {}
    for f in FILES:
        shutil.copyfile(os.path.join(SRC, f), os.path.join(DST, f))
        manifest[f] = hashlib.sha256(open(os.path.join(DST, f), "rb").read()).hexdigest()
    open(os.path.join(DST, "__init__.py"), "w").close()
    json.dump({"source": SRC, "sha256": manifest}, open(os.path.join(os.path.dirname(DST), "MANIFEST.json"), "w"), indent=1)
    return DST


if __name__ == "__main__":
    print(build() or f"{SRC} not present: nothing vendored")
