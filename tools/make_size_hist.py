"""Derive heavy-atom-count histograms from the reference's CSV files.

Run ONCE in the build container (the CSVs live under /root/reference/Data, which does
not exist on the GPU box).  Output: eagcn_b200/size_hist.json -- {dataset: {n_atoms: count}}.
Only the size distribution is kept; no molecule data is copied.
The crude SMILES tokenizer is the one SURVEY.md 8(d) specifies; it reproduces the
maxima the reference itself quotes (utils.py:524,590: 132 tox21 / 222 hiv / 115 lipo / 24 freesolv).
"""
import csv, json, re, sys, collections

TOK = re.compile(r"\[[^\]]+\]|Cl|Br|[BCNOPSFI]|[bcnops]")
FILES = {"tox21": ("tox21.csv", "smiles"), "hiv": ("HIV.csv", "smiles"),
         "lipo": ("Lipophilicity.csv", "smiles"), "freesolv": ("SAMPL.csv", "smiles")}

def main(root="/root/reference/Data", out="eagcn_b200/size_hist.json"):
    res = {}
    for name, (fn, col) in FILES.items():
        hist = collections.Counter()
        with open(f"{root}/{fn}", newline="") as f:
            for row in csv.DictReader(f):
                s = (row.get(col) or "").strip()
                if not s:
                    continue
                n = len(TOK.findall(s))
                if n >= 2:
                    hist[n] += 1
        res[name] = {str(k): hist[k] for k in sorted(hist)}
        tot = sum(hist.values()); mean = sum(k * v for k, v in hist.items()) / tot
        print(name, "mols", tot, "mean %.2f" % mean, "max", max(hist), file=sys.stderr)
    with open(out, "w") as f:
        json.dump(res, f, separators=(",", ":"))

if __name__ == "__main__":
    main(*sys.argv[1:])
