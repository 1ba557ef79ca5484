"""Recipe for oracle/_ref: make the UNMODIFIED reference modules of the hot path available next to the oracle so that
they travel to the GPU box (oracle/_ref/ is git-ignored, not gpurun-ignored — it never enters the history).

    python oracle/build_ref.py          (also run by __graft_entry__.build() when /root/reference is present)

Files staged (byte-identical, verified by hash after the copy) — the SURVEY.md §8(a) set plus what they import:
    models/*.py  losses.py  utils/train_util.py
They are used ONLY as the CPU / torch.cuda baseline (`bench.py --impl reference`, `cpu_baseline.kind = "reference"`,
`gpu_baseline`) and to generate fixtures; nothing under texttoaudiogrounding_b200/ imports them.
TEST / MEASUREMENT INFRASTRUCTURE ONLY."""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("TAG_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = ["losses.py", "utils/train_util.py"] + [
    f"models/{f}" for f in ("__init__.py", "align.py", "audio_encoder.py", "audio_text_model.py", "base.py",
                            "cross_encoder.py", "hf_modeling_grounding.py", "match.py", "panns.py", "sim_pooling.py",
                            "text_encoder.py", "utils.py")]


def _sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def build(verbose: bool = True) -> bool:
    """Returns True when oracle/_ref holds the reference files (copied now or already there)."""
    if not os.path.isdir(os.path.join(SRC, "models")):
        ok = os.path.isfile(os.path.join(DST, "models", "audio_encoder.py"))
        if verbose:
            print(f"build_ref: {SRC} not present; oracle/_ref {'already staged' if ok else 'absent'}")
        return ok
    manifest = []
    for rel in FILES:
        src, dst = os.path.join(SRC, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        assert _sha(src) == _sha(dst), rel
        manifest.append(f"{_sha(dst)}  {rel}")
    with open(os.path.join(DST, "MANIFEST.sha256"), "w") as f:
        f.write("\n".join(manifest) + "\n")
    if verbose:
        print(f"build_ref: staged {len(FILES)} unmodified reference files under {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
