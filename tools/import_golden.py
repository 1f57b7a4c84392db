#!/usr/bin/env python3
"""Copy the reference's own state fixtures (DATA, not source) into tests/golden/.

The GPU box has no /root/reference, so every fixture a test needs must travel
with the repo.  Files are copied byte-for-byte; MANIFEST.json records where
each one came from (path under /root/reference/test) and its sha256, so the
copy can be re-verified at any time with `--check`.

Run in the build container only:   python tools/import_golden.py
"""
import hashlib
import json
import os
import shutil
import sys

REF = "/root/reference/test"
DST = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# golden name -> path under /root/reference/test
FILES = {
    # KAT-1: the tc-smoke pair (tc-smoke.cmd:11-15): 1010 bodies, 20 steps
    "kat1_start.state": "regression/_baseline-2024-09-18/1000_bodies-START.state",
    "kat1_end.state": "regression/tc-smoke/REFERENCE-END.state",
    # KAT-2: same bodies, header says interactions = 0
    "kat2_start.state": "regression/_baseline-d5de5369/1000_bodies-START.state",
    # KAT-3: 501 bodies, 500 steps (statistical only; zstd framed)
    "kat3_start.state": "regression/_baseline-ea39db36/500_bodies-START.state",
    "kat3_end.state": "regression/tc-500_bodies+500_cycles/REFERENCE-END.state",
    # start-state corpus (no END shipped)
    "realg_test.state": "user/session/realg-test",
    "multi_cluster_640.save": "user/session/640 bodies multi-cluster big world.save",
    "two_bodies_realg.state": "user/session/2bodies-realg",
    "two_upclose.state": "user/session/2upclose",
    "earth_moon.state": "user/session/Earth-Moon",
    "one_body_realg_double.save": "user/session/1body-realg [double].save",
    "uncompressed_0.0.1.sav": "user/session/test-uncompressed-0.0.1.sav",
    "dense_black_hole.save": "user/session/gems/dense black hole.save",
    "snapshot_4_14k.save": ".save/snapshot_4.save",
    "two_moons.save": ".save/snapshot-2_moons.save",
}


def sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def main():
    check = "--check" in sys.argv
    os.makedirs(DST, exist_ok=True)
    manifest = {}
    for name, rel in FILES.items():
        src = os.path.join(REF, rel)
        dst = os.path.join(DST, name)
        if check:
            assert sha(src) == sha(dst), name
        else:
            shutil.copyfile(src, dst)
            os.chmod(dst, 0o644)
        manifest[name] = {"source": "test/" + rel, "sha256": sha(dst), "bytes": os.path.getsize(dst)}
    if not check:
        with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
            json.dump(manifest, f, indent=1, sort_keys=True)
    print("ok", len(manifest), "files")


if __name__ == "__main__":
    main()
