"""Instruction mix of every kernel of libgpis_b200.so from its SASS (cuobjdump -sass), plus an excerpt of the hot
loops of the evaluation and training kernels: what proves packed FFMA2, TMA bulk copies (UBLKCP), cp.async (LDGSTS)
and the absence of tensor-core instructions. Writes profiles/<tag>_sass_summary.txt."""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
lib = os.path.join(ROOT, "gpismap_b200", "libgpis_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
arch = re.findall(r"arch = (sm_\w+)", sass)
cur, kern = None, {}
for ln in sass.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0].replace("gpis::", "").replace("void ", "")
        kern[cur] = {"ops": {}, "lines": []}
        continue
    if cur:
        m = re.search(r"^\s+/\*([0-9a-f]+)\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", ln)
        if m:
            op = m.group(2)
            kern[cur]["ops"][op] = kern[cur]["ops"].get(op, 0) + 1
            kern[cur]["lines"].append(ln.rstrip())
keys = ["FFMA2", "FFMA", "FMUL", "FADD", "DFMA", "MUFU", "LDS", "STS", "LDG", "STG", "LDGSTS", "UBLKCP", "SYNCS", "SHFL", "BAR", "ATOM", "ATOMS", "RED", "HMMA", "UTCMMA", "LDL", "STL"]
out = [f"# cuobjdump -sass gpismap_b200/libgpis_b200.so — architectures: {sorted(set(arch))}",
       "# instruction counts per kernel (static): FFMA2 = packed fp32x2 FMA, UBLKCP = TMA bulk copy, LDGSTS = cp.async, SYNCS = mbarrier ops;",
       "# HMMA / UTCMMA (tensor cores) must be absent: independent fp32 factorizations and solves with a 1e-4 contract (north_star).",
       f"{'kernel':34s} {'total':>7s} " + " ".join(f"{k:>6s}" for k in keys)]
for k, v in sorted(kern.items(), key=lambda kv: -sum(kv[1]["ops"].values())):
    out.append(f"{k[:34]:34s} {sum(v['ops'].values()):7d} " + " ".join(f"{v['ops'].get(x, 0):6d}" for x in keys))
tot = {x: sum(v["ops"].get(x, 0) for v in kern.values()) for x in keys}
out.append(f"{'all kernels':34s} {sum(sum(v['ops'].values()) for v in kern.values()):7d} " + " ".join(f"{tot[x]:6d}" for x in keys))


def excerpt(name, pattern, before=6, after=22):
    for k, v in kern.items():
        if k.startswith(name):
            idx = [i for i, l in enumerate(v["lines"]) if pattern in l]
            if not idx:
                continue
            # the densest window of `pattern`
            best = max(idx, key=lambda i: sum(1 for j in idx if i <= j < i + after))
            out.append("")
            out.append(f"## {k}: hot loop excerpt (densest {pattern} window)")
            out.extend(l[:150] for l in v["lines"][max(0, best - before):best + after])
            return


excerpt("k_eval_v3<8>", "FFMA2")
excerpt("k_leaf_train", "FFMA2")
excerpt("k_leaf_train", "UBLKCP", 3, 6)
excerpt("k_eval_v3<8>", "LDGSTS", 3, 8)
open(os.path.join(ROOT, "profiles", f"{tag}_sass_summary.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out[:40]))
