"""Copies the reference's shipped weight files (parity fixtures) and its test-driver scripts from
/root/reference into tests/golden/_ref_data/ (git-ignored; travels to the GPU box with the snapshot).
Run by __graft_entry__.build() in the build container; a no-op where /root/reference is absent."""
import os
import shutil
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("DVMVS_REFERENCE_ROOT", "/root/reference")


def fetch(verbose=True):
    if not os.path.isdir(os.path.join(REF, "dvmvs")):
        if verbose:
            print("fetch_fixtures: no reference tree at", REF, "- nothing fetched")
        return False
    for net in ("fusionnet", "pairnet"):
        src = os.path.join(REF, "dvmvs", net, "weights")
        dst = os.path.join(REPO, "tests", "golden", "_ref_data", "weights", net)
        os.makedirs(dst, exist_ok=True)
        for f in sorted(os.listdir(src)):
            s, d = os.path.join(src, f), os.path.join(dst, f)
            if not os.path.isfile(d) or os.path.getsize(d) != os.path.getsize(s):
                shutil.copyfile(s, d)
                if verbose:
                    print("fetched", net, f)
    # the reference's test drivers, verbatim, for the "runs unchanged" test (tests/test_gpu_reference_script.py): git-ignored like
    # the weights, executed from there on the GPU box, never imported by the product
    for net in ("fusionnet", "pairnet"):
        dst = os.path.join(REPO, "tests", "golden", "_ref_data", "scripts", net)
        os.makedirs(dst, exist_ok=True)
        for f in ("run-testing.py", "run-testing-online.py"):
            s = os.path.join(REF, "dvmvs", net, f)
            if os.path.isfile(s):
                shutil.copyfile(s, os.path.join(dst, f))
                if verbose:
                    print("fetched script", net, f)
    return True


if __name__ == "__main__":
    sys.exit(0 if fetch() else 0)
