#!/usr/bin/env python
"""Adds the large goldens to tests/golden/golden.json (run in the build container):

  --ranks   64 MiB text blocks for seeds 0x5EED0064 + 1 .. + 7 (the files bench.py's ranks 1..7
            compress), each through the UNMODIFIED reference binary (oracle/_ref/ref_compress, ~1 min
            each, run in parallel)                                             "source": "reference"
  --block1g the 1 GiB text block of BASELINE config 5(ii), seed 0x5EED1024, through the pinned
            oracle restatement (oracle/bzap_oracle.c, ~40 GB of RAM, O(N log N)); the reference
            needs > 16 min and 14 GiB for it                        "source": "oracle"
  --block1g-ref  the same block through the reference binary too (optional, slow); records whether
            the two agree.
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O  # noqa: E402
import workloads as W  # noqa: E402

OUT = os.path.join(HERE, "golden.json")
BASE_SEED = 0x5EED0064
SEED_1G = 0x5EED1024


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.uint8).tobytes()).hexdigest()


def entry(data, blob, source):
    primary, n, tree, payload = O.split_container(blob)
    return {"n": int(n), "primary": int(primary), "tree_hex": tree.tobytes().hex(), "payload_bytes": int(payload.size),
            "total": int(blob.size), "sha256": sha(blob), "input_sha256": sha(data), "source": source}


def update(section, key, e):
    G = json.load(open(OUT))
    G.setdefault(section, {})[key] = e
    json.dump(G, open(OUT, "w"), indent=1, sort_keys=True)


def ref_compress_file(data):
    with tempfile.TemporaryDirectory() as tmp:
        p = os.path.join(tmp, "in")
        data.tofile(p)
        subprocess.run([os.path.join(O.REF_DIR, "ref_compress"), p, p + ".bz"], check=True, stdout=subprocess.DEVNULL)
        return np.fromfile(p + ".bz", dtype=np.uint8)


def main():
    O.build_oracle()
    if "--ranks" in sys.argv:
        assert O.have_ref()

        def one(r):
            d = W.synthetic_text(1 << 26, BASE_SEED + r)
            t0 = time.time()
            blob = ref_compress_file(d)
            print("rank seed +%d: reference %.1fs -> %d" % (r, time.time() - t0, blob.size), flush=True)
            return r, entry(d, blob, "reference")
        with ThreadPoolExecutor(7) as ex:
            for r, e in ex.map(one, range(1, 8)):
                e["seed"] = "0x5EED0064+%d" % r
                update("text_rank", str(r), e)
    if "--block1g" in sys.argv:
        d = W.synthetic_text(1 << 30, SEED_1G)
        print("generated 1 GiB text, sha", sha(d)[:16], flush=True)
        t0 = time.time()
        blob = O.o_compress(d)
        print("oracle 1 GiB compress %.1fs -> %d" % (time.time() - t0, blob.size), flush=True)
        e = entry(d, blob, "oracle")
        e["seed"] = "0x5EED1024"
        update("text", str(1 << 30), e)
    if "--block1g-ref" in sys.argv:
        assert O.have_ref()
        d = W.synthetic_text(1 << 30, SEED_1G)
        t0 = time.time()
        blob = ref_compress_file(d)
        print("reference 1 GiB compress %.1fs -> %d" % (time.time() - t0, blob.size), flush=True)
        G = json.load(open(OUT))
        e = G["text"].get(str(1 << 30))
        s = sha(blob)
        if e is None:
            e = entry(d, blob, "reference")
            e["seed"] = "0x5EED1024"
        else:
            e["reference_sha256"] = s
            e["oracle_matches_reference"] = bool(s == e["sha256"])
            if s == e["sha256"]:
                e["source"] = "oracle, confirmed by the reference binary"
        update("text", str(1 << 30), e)
        print("reference agrees with oracle:", e.get("oracle_matches_reference"), flush=True)


if __name__ == "__main__":
    main()
