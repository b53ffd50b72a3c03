"""Regenerates the fixtures in tests/golden/ from the reference checkout.

Run in the build container only (needs /root/reference, which does not exist on
the GPU box):  python tests/golden/make_golden_from_reference.py

Sources (reference-produced data, not code):
  runs/*.txt -- 20 search logs over the same 1024 5x5 roots; every line lists the
  root's children IN CHILD ORDER as action:visits:eval:std_dev[:logit]
  (python/improved_policy.py:29-33).
Outputs:
  runs_moves_5x5.txt.gz   child-order move lists of the 1024 roots (one file's
                          worth; the generator asserts all 20 files agree)
  runs_visits.json        per sequential-halving log: k, budget-derived most common visit
                          multiset over its 1024 lines, and the first puct.txt line in full
"""
import collections
import gzip
import json
import os
import re

SRC = "/root/reference/runs"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    files = sorted(os.listdir(SRC))
    moves_ref = None
    visits = {}
    for f in files:
        lines = open(os.path.join(SRC, f)).read().strip().split("\n")
        assert len(lines) == 1024, (f, len(lines))
        moves = [",".join(t.split(":")[0] for t in ln.strip(",").split(",")) for ln in lines]
        if moves_ref is None:
            moves_ref = moves
        else:
            assert moves == moves_ref, f"{f}: roots differ"
        m = re.match(r"seqhal_(\d+)_", f)
        if m:
            multisets = collections.Counter()
            for ln in lines:
                vs = sorted((int(t.split(":")[1]) for t in ln.strip(",").split(",")), reverse=True)
                multisets[json.dumps([v for v in vs if v > 0])] += 1
            top, count = multisets.most_common(1)[0]
            visits[f] = {"k": int(m.group(1)), "top_multiset": json.loads(top), "lines_with_it": count,
                         "distinct_multisets": len(multisets)}
    with gzip.open(os.path.join(HERE, "runs_moves_5x5.txt.gz"), "wt", compresslevel=9) as fh:
        fh.write("\n".join(moves_ref) + "\n")
    first_puct = open(os.path.join(SRC, "puct.txt")).readline().strip()
    with open(os.path.join(HERE, "runs_visits.json"), "w") as fh:
        json.dump({"seqhal": visits, "puct_first_line": first_puct}, fh, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
