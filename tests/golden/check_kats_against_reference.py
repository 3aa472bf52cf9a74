"""Re-check tests/golden/reference_kats.json against the mounted reference tree.

Run here (the build container) only: `python tests/golden/check_kats_against_reference.py`.
For every entry it opens the cited file and asserts that the literals (sequence,
k-mer strings, hash constants) appear in the cited line range.  The GPU box has no
/root/reference; nothing in tests/ depends on this script at run time.
"""
import json
import os
import re
import sys

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def lines(cite):
    m = re.match(r"([\w./]+):(\d+)(?:-(\d+))?", cite)
    path, a, b = m.group(1), int(m.group(2)), int(m.group(3) or m.group(2))
    with open(os.path.join(REF, path)) as f:
        src = f.read().split("\n")
    return "\n".join(src[max(0, a - 3): b + 2])


def main():
    if not os.path.isdir(REF):
        print("reference tree not mounted; nothing to check")
        return 0
    kats = json.load(open(os.path.join(HERE, "reference_kats.json")))
    n = 0

    def need(cite, *lits):
        nonlocal n
        txt = lines(cite)
        for lit in lits:
            if lit == "":
                continue
            # the reference writes RNA with U and some literals lower-case
            if lit not in txt and lit.replace("T", "U") not in txt and lit.lower() not in txt.lower():
                raise SystemExit(f"literal {lit!r} not found at {cite}")
            n += 1

    for e in kats["fx_hash"]:
        need(e["cite"], e["kmer"], e["hash"])
    for e in kats["as_integer"]:
        need(e["cite"], e["kmer"], e["value"])
    for e in kats["canonical"]:
        need(e["cite"], e["seq"] if e["k"] == 3 else e["expect"][0], *e["expect"])
    for e in kats["fwrv"]:
        need(e["cite"], e["seq"], *[x for p in e["expect"] for x in p])
    for e in kats["unambiguous"]:
        need(e["cite"], e["seq"], *[p[0].replace("T", "U") for p in e["expect"]])
    for e in kats["unambiguous_starts"]:
        need(e["cite"], e["seq"])
    for e in kats["shift_from_4to2"]:
        need(e["cite"], e["kmer"], e["seq"], e["expect"])
    for e in kats["strict_4to2_errors"]:
        need(e["cite"], e["seq"])
    for e in kats["iscanonical"]:
        need(e["cite"], e["kmer"])
    for key, e in kats["differential_sequences"].items():
        if key.startswith("_"):
            continue
        need(e["cite"], *e["seqs"])
    print(f"{n} literals verified against {REF}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
