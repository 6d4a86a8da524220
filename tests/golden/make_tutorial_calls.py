"""Extracts every call of the hot-path API from the reference's driver scripts (tutorials/*.jl, examples/*.jl) into
tests/golden/tutorial_calls.json: function name, number of positional arguments, keyword names.  tests/test_julia_shim_signatures.py
checks each call against the method signatures of julia/SmoQyElPhB200.jl (the drop-in boundary).
Run in the build container (needs /root/reference): python tests/golden/make_tutorial_calls.py"""
import glob
import json
import os
import re

API = ["SymFermionDetMatrix", "AsymFermionDetMatrix", "KPMPreconditioner", "PFFCalculator", "EFAPFFHMCUpdater", "GreensEstimator",
       "hmc_update!", "reflection_update!", "swap_update!", "radial_update!", "make_measurements!", "update_chemical_potential!"]
REF = os.environ.get("SQ_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def strip_comments(src):
    out = []
    for line in src.splitlines():
        q = False
        cut = len(line)
        for i, ch in enumerate(line):
            if ch == '"':
                q = not q
            elif ch == "#" and not q:
                cut = i
                break
        out.append(line[:cut])
    return "\n".join(out)


def split_top(text, seps=","):
    """split on top-level separators (outside brackets and strings)"""
    parts, depth, q, cur = [], 0, False, ""
    for ch in text:
        if ch == '"':
            q = not q
        if not q:
            if ch in "([{":
                depth += 1
            elif ch in ")]}":
                depth -= 1
            elif ch in seps and depth == 0:
                parts.append((cur, ch))
                cur = ""
                continue
        cur += ch
    parts.append((cur, ""))
    return parts


def parse_args(argtext):
    """-> (number of positional arguments, sorted keyword names)"""
    npos, kws, after_semi = 0, [], False
    for piece, sep in split_top(argtext, ",;"):
        p = piece.strip()
        if p:
            m = re.match(r"^([A-Za-z_Ͱ-Ͽ′][\wͰ-Ͽ′!]*)\s*=(?!=)", p)
            if m:
                kws.append(m.group(1))
            elif after_semi:
                kws.append(p)                     # `; rng` shorthand for rng = rng
            else:
                npos += 1
        if sep == ";":
            after_semi = True
    return npos, sorted(kws)


def calls_in(path):
    src = strip_comments(open(path, encoding="utf-8").read())
    found = []
    for name in API:
        for m in re.finditer(r"(?<![\w.!])" + re.escape(name) + r"\(", src):
            i, depth = m.end(), 1
            while depth and i < len(src):
                depth += {"(": 1, ")": -1}.get(src[i], 0)
                i += 1
            npos, kws = parse_args(src[m.end():i - 1])
            found.append({"function": name, "positional": npos, "keywords": kws, "line": src.count("\n", 0, m.start()) + 1})
    return found


def main():
    out = {}
    for path in sorted(glob.glob(os.path.join(REF, "tutorials", "*.jl")) + glob.glob(os.path.join(REF, "examples", "*.jl"))):
        out[os.path.relpath(path, REF)] = calls_in(path)
    with open(os.path.join(HERE, "tutorial_calls.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    n = sum(len(v) for v in out.values())
    print(f"{n} calls from {len(out)} driver scripts")


if __name__ == "__main__":
    main()
