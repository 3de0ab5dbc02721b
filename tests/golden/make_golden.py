#!/usr/bin/env python
"""Regenerates tests/golden/golden.json from the UNMODIFIED reference (oracle/_ref, built by
oracle/Makefile from /root/reference) and, where the reference's O(N^2 log N) sort cannot
finish, from the pinned oracle restatement (entries say which: "source": "reference"|"oracle").

Run in the build container (needs /root/reference):  python tests/golden/make_golden.py [--big]
--big also produces the 64 MiB text golden (reference: ~1 min compress) and the 16 MiB
degenerate goldens (oracle).  Hashes are only meaningful for this image's glibc.
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O  # noqa: E402
import workloads as W  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.uint8).tobytes()).hexdigest()


def entry(data, blob, source, full=False):
    primary, n, tree, payload = O.split_container(blob)
    e = {"n": int(n), "primary": int(primary), "tree_hex": tree.tobytes().hex(), "payload_bytes": int(payload.size),
         "total": int(blob.size), "sha256": sha(blob), "input_sha256": sha(data), "source": source}
    if full:
        e["file_hex"] = blob.tobytes().hex()
    return e


def main():
    big = "--big" in sys.argv
    out_path = os.path.join(HERE, "golden.json")
    G = json.load(open(out_path)) if os.path.exists(out_path) else {}
    O.build_oracle()
    assert O.have_ref(), "build oracle/_ref first (make -C oracle)"

    cal = W.calgary()
    G["calgary"] = {}
    for name in W.CALGARY_FILES:
        d = np.frombuffer(cal[name], dtype=np.uint8)
        blob = O.ref_compress(d)
        assert np.array_equal(O.o_compress(d), blob), name
        G["calgary"][name] = entry(d, blob, "reference")
        print("calgary", name, G["calgary"][name]["total"])

    # small known-answer vectors: full files in hex.  N is tiny, so these sit
    # outside App. B's validity windows; only inputs with no frequency ties are tie-safe, the
    # rest are still what the reference emits here and what the oracle reproduces.
    kats = {"a": b"a", "aa": b"aa", "abc": b"abc", "a1000": b"a" * 1000, "zeros1000": b"\0" * 1000,
            "banana": b"banana", "abracadabra": b"abracadabra" * 3,
            "mississippi": b"mississippi$" * 5}
    G["kat"] = {}
    for k, v in kats.items():
        d = np.frombuffer(v, dtype=np.uint8)
        blob = O.ref_compress(d)
        G["kat"][k] = entry(d, blob, "reference", full=True)
        G["kat"][k]["input_hex"] = v.hex()
        G["kat"][k]["oracle_matches_reference"] = bool(np.array_equal(O.o_compress(d), blob))
        print("kat", k, G["kat"][k]["total"], G["kat"][k]["oracle_matches_reference"])

    # BWT-only KATs: primary index of periodic inputs, via the stage harness
    bw = {"ba*1024": b"ba" * 1024, "abcabc": b"abcabc", "aaaa": b"aaaa", "zzzzy*50": b"zzzzy" * 50,
          "cab*300": b"cab" * 300, "banana*40+x": b"banana" * 40 + b"x", "b+ab*100": b"b" + b"ab" * 100,
          "dcba*2000[:4098]": (b"dcba" * 2000)[:4098]}
    G["bwt_kat"] = {}
    for k, v in bw.items():
        last, p = O.r_bwt(v)
        ol, op = O.o_bwt(v)
        assert p == op and np.array_equal(last, ol), k
        G["bwt_kat"][k] = {"input_hex": v.hex(), "primary": p, "last_sha256": sha(last), "source": "reference"}
        print("bwt_kat", k, p)

    # degenerate inputs at N = 16384 (inside an App. B validity window; reference needs <=3 s each)
    G["degenerate_16k"] = {}
    for kind in W.DEGENERATE_KINDS:
        d = W.degenerate(kind, 16384)
        blob = O.ref_compress(d)
        ob = O.o_compress(d)
        G["degenerate_16k"][kind] = entry(d, blob, "reference")
        G["degenerate_16k"][kind]["oracle_matches_reference"] = bool(np.array_equal(ob, blob))
        print("deg16k", kind, blob.size, G["degenerate_16k"][kind]["oracle_matches_reference"])

    # synthetic text (C3 generator) at 1 MiB and 4 MiB through the real reference
    G.setdefault("text", {})
    for n in (1 << 20, 1 << 22):
        d = W.synthetic_text(n)
        blob = O.ref_compress(d)
        assert np.array_equal(O.o_compress(d), blob)
        G["text"][str(n)] = entry(d, blob, "reference")
        print("text", n, blob.size)

    if big:
        n = 1 << 26
        d = W.synthetic_text(n)
        t0 = time.time()
        blob = O.ref_compress(d)
        print("text 64MiB reference compress %.1fs -> %d" % (time.time() - t0, blob.size))
        G["text"][str(n)] = entry(d, blob, "reference")
        json.dump(G, open(out_path, "w"), indent=1, sort_keys=True)
        t0 = time.time()
        ob = O.o_compress(d)
        print("text 64MiB oracle compress %.1fs match=%s" % (time.time() - t0, np.array_equal(ob, blob)))
        G["text"][str(n)]["oracle_matches_reference"] = bool(np.array_equal(ob, blob))
        G["degenerate_16m"] = {}
        for kind in W.DEGENERATE_KINDS:
            d = W.degenerate(kind, 1 << 24)
            t0 = time.time()
            ob = O.o_compress(d)
            G["degenerate_16m"][kind] = entry(d, ob, "oracle")
            print("deg16m", kind, ob.size, "%.1fs" % (time.time() - t0))
            json.dump(G, open(out_path, "w"), indent=1, sort_keys=True)

    json.dump(G, open(out_path, "w"), indent=1, sort_keys=True)
    print("wrote", out_path)


if __name__ == "__main__":
    main()
