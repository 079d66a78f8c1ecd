"""Join an ncu SASS source page with nvdisasm -gi line info: stall samples per source line and
per enclosing function.

    ncu -i X.ncu-rep --page source --csv > sass.csv
    cuobjdump -xelf all zf_batched.o; nvdisasm -gi zf_batched.sm_100a.cubin > lines.txt
    python profiles/stall_by_line.py sass.csv lines.txt '<mangled kernel name>' [top]
"""
import collections
import csv
import re
import sys


def load_lines(path, kernel):
    """offset -> [(file, line), ...] innermost first"""
    out = {}
    active = False
    chain = []
    pending = []
    with open(path) as fh:
        for ln in fh:
            if ln.startswith(".text."):
                active = ln.strip() == f".text.{kernel}:"
                continue
            if not active:
                continue
            m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', ln)
            if m:
                locs = [(m.group(1), int(m.group(2)))]
                for mm in re.finditer(r'inlined at "([^"]+)", line (\d+)', m.group(3)):
                    locs.append((mm.group(1), int(mm.group(2))))
                pending.append(locs)
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,6})\*/", ln)
            if m:
                if pending:
                    # first comment = innermost location with its inline chain; following plain
                    # lines continue the chain outwards
                    chain = [loc for locs in pending for loc in locs]
                    pending = []
                out[int(m.group(1), 16)] = chain
    return out


def function_table(files):
    tabs = {}
    pat = re.compile(r"^\s{0,2}(?:template\s*<[^>]*>\s*)?(?:__device__|__global__|static|inline|__host__)"
                     r".*?([A-Za-z_][A-Za-z_0-9:<>, ]*)\s*\(")
    for f in files:
        starts = []
        try:
            src = open(f).read().split("\n")
        except OSError:
            continue
        for i, ln in enumerate(src, 1):
            if ln.startswith("    "):
                continue
            m = pat.match(ln)
            if m and "=" not in ln.split("(")[0] and not ln.strip().startswith("//"):
                name = ln.split("(")[0].split()[-1]
                starts.append((i, name))
        tabs[f] = starts
    return tabs


def func_of(tabs, f, line):
    best = "?"
    for s, name in tabs.get(f, []):
        if s <= line + 3:
            best = name
        else:
            break
    return best


def main():
    sass, lines, kernel = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    loc = load_lines(lines, kernel)
    rows = list(csv.reader(open(sass)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    ia, isamp, iinst = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    base = None
    by_line = collections.Counter()
    inst_line = collections.Counter()
    by_func = collections.Counter()
    inst_func = collections.Counter()
    by_outer = collections.Counter()
    stalls = collections.Counter()
    files = set()
    recs = []
    for r in rows[hi + 1:]:
        if len(r) <= isamp or not r[ia].startswith("0x"):
            continue
        a = int(r[ia], 16)
        if base is None:
            base = a
        chain = loc.get(a - base, [])
        smp, ins = float(r[isamp] or 0), float(r[iinst] or 0)
        for i, h in stall_cols:
            stalls[h] += float(r[i] or 0)
        recs.append((chain, smp, ins))
        for f, _ in chain:
            files.add(f)
    tabs = function_table(files)
    tot = sum(s for _, s, _ in recs)
    toti = sum(i for _, _, i in recs)
    for chain, smp, ins in recs:
        if not chain:
            key = ("?", 0)
            fn = outer = "?"
        else:
            key = chain[0]
            fn = func_of(tabs, *chain[0])
            # outermost-but-one frame: the call site inside the kernel's own body
            outer = " <- ".join(func_of(tabs, *c) for c in chain[:4])
        by_line[key] += smp
        inst_line[key] += ins
        by_func[fn] += smp
        inst_func[fn] += ins
        by_outer[outer] += smp
    print(f"total samples {tot:.0f}, warp instructions {toti:.0f}")
    print("stall reasons:", ", ".join(f"{h[6:]} {v / tot * 100:.1f}%" for h, v in stalls.most_common(8)))
    print("\n-- by innermost function (samples %, instructions %)")
    for fn, v in by_func.most_common(25):
        print(f"  {v / tot * 100:5.1f}%  {inst_func[fn] / toti * 100:5.1f}%  {fn}")
    print("\n-- by inline chain")
    for fn, v in by_outer.most_common(30):
        print(f"  {v / tot * 100:5.1f}%  {fn}")
    print("\n-- by source line")
    for (f, l), v in by_line.most_common(top):
        src = ""
        try:
            src = open(f).read().split("\n")[l - 1].strip()[:90]
        except (OSError, IndexError):
            pass
        print(f"  {v / tot * 100:5.1f}%  {inst_line[(f, l)] / toti * 100:5.1f}%  {f.split('/')[-1]}:{l}  {src}")


if __name__ == "__main__":
    main()
