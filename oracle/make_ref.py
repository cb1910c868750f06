"""TEST / BASELINE INFRASTRUCTURE (not product code): place the UNMODIFIED reference modules of the hot path under
oracle/_ref/ so that the reference arm of bench.py (`--impl reference`, and the `gpu_eager` comparator) runs the
reference's own code on the GPU box, where /root/reference is not mounted.

    python oracle/make_ref.py          # copies /root/reference/finetune/{losses,optimizers,config}.py -> oracle/_ref/

oracle/_ref/ is git-ignored (reference sources never enter the history) but NOT gpurun-ignored, so it travels with
the snapshot like the built libcfa_b200.so.  __graft_entry__.build() calls this when /root/reference is present.
Only bench.py's reference legs and tests/ import from oracle/_ref.
"""
import hashlib
import os
import shutil
import sys

SRC = "/root/reference/finetune"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
FILES = ("losses.py", "optimizers.py", "config.py")


def make_ref(verbose: bool = True) -> bool:
    if not os.path.isdir(SRC):
        if verbose:
            print(f"[make_ref] {SRC} not mounted: oracle/_ref left as is", file=sys.stderr)
        return os.path.isdir(DST)
    os.makedirs(DST, exist_ok=True)
    for f in FILES:
        shutil.copyfile(os.path.join(SRC, f), os.path.join(DST, f))
        if verbose:
            h = hashlib.sha256(open(os.path.join(DST, f), "rb").read()).hexdigest()[:16]
            print(f"[make_ref] {f} sha256/16 = {h}", file=sys.stderr)
    return True


def import_ref():
    """(losses, optimizers) modules of the unmodified reference, or None when oracle/_ref has not been made."""
    if not all(os.path.exists(os.path.join(DST, f)) for f in FILES):
        return None
    sys.path.insert(0, DST)
    try:
        for name in ("config", "losses", "optimizers"):
            sys.modules.pop(name, None)
        import losses as ref_losses          # bare module names, as the reference imports them (finetuner.py:14-16)
        import optimizers as ref_opt
        return ref_losses, ref_opt
    except Exception:
        return None
    finally:
        sys.path.remove(DST)


if __name__ == "__main__":
    ok = make_ref()
    print("oracle/_ref ready" if ok else "oracle/_ref missing")
