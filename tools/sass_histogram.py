"""SASS opcode histogram of every kernel in libnunchaku_b200.so (cuobjdump -sass on the object files) -> profiles/<tag>_sass_histogram.md.
Shows at a glance which hardware paths a kernel uses: UTCQMMA / UTCHMMA / UTCOMMA (tcgen05.mma kinds), UTCCP (tcgen05.cp), LDTM (tcgen05.ld),
UTMALDG / UTMASTG / UBLKCP (TMA), SYNCS (mbarrier), HMMA (legacy mma.sync), MUFU, and how much plain ALU work surrounds them.

    python tools/sass_histogram.py [--tag r02]
"""
import argparse
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEY = ["UTCQMMA", "UTCOMMA", "UTCHMMA", "UTCIMMA", "UTCMXQMMA", "UTCCP", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF", "SYNCS", "HMMA", "IMMA", "LDSM", "MUFU",
       "ACQBULK", "ELECT", "MEMBAR", "FENCE", "ATOM", "RED", "LDG", "STG", "LDS", "STS", "BAR"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tag", default="r02")
    args = ap.parse_args()
    out = [f"# SASS opcode histogram per kernel ({args.tag})", "",
           "`cuobjdump -sass nunchaku_b200/_lib/obj/*.o`, instructions counted statically (one template instantiation per kernel family is listed:",
           "the bf16 one with the most instructions).  tcgen05.mma appears as UTC*MMA, tcgen05.cp as UTCCP, tcgen05.ld as LDTM, TMA as UTMALDG / UTMASTG /",
           "UBLKCP, mbarriers as SYNCS; HMMA is the legacy mma.sync path (only the quantizer's 32-rank projection uses it).", ""]
    for obj in sorted(glob.glob(os.path.join(ROOT, "nunchaku_b200", "_lib", "obj", "*.o"))):
        txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
        kernels = collections.OrderedDict()
        cur = None
        for line in txt.splitlines():
            m = re.match(r"\s*Function : (\S+)", line)
            if m:
                cur = m.group(1)
                kernels[cur] = collections.Counter()
                continue
            m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
            if m and cur:
                kernels[cur][m.group(1)] += 1
        if not kernels:
            continue
        demangled = {}
        names = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
        for k, d in zip(kernels, names):
            d = re.sub(r"\(anonymous namespace\)::|nb200::|ptx::", "", d).replace("void ", "")
            depth, cut = 0, len(d)
            for i, ch in enumerate(d):   # the argument list starts at the first '(' outside the template brackets
                if ch == "<":
                    depth += 1
                elif ch == ">":
                    depth -= 1
                elif ch == "(" and depth == 0:
                    cut = i
                    break
            demangled[k] = d[:cut]
        # one instantiation per family: prefer bf16, then the largest
        fam = {}
        for k, c in kernels.items():
            name = demangled[k]
            base = name.split("<")[0]
            if base == "gemm_w4a4_kernel":   # one line per epilogue (last template argument: 0 default, 1 fused quantise, 2 RMSNorm + RoPE, 3 LiteLA)
                base += " epilogue " + name.rstrip(">").split(",")[-1].strip()
            score = (("__nv_bfloat16" in name), sum(c.values()))
            if base not in fam or score > fam[base][0]:
                fam[base] = (score, k)
        out.append(f"## {os.path.basename(obj)}")
        out.append("")
        for base, (_, k) in fam.items():
            c = kernels[k]
            total = sum(c.values())
            keys = [f"{op} {c[op]}" for op in KEY if c.get(op)]
            # fold per-opcode variants (e.g. UTCQMMA.2CTA) by prefix
            pref = collections.Counter()
            for op, n in c.items():
                for kk in KEY:
                    if op.startswith(kk):
                        pref[kk] += n
                        break
            keys = [f"{kk} {pref[kk]}" for kk in KEY if pref.get(kk)]
            top = ", ".join(f"{op} {n}" for op, n in c.most_common(8))
            out.append(f"* `{demangled[k]}` — {total} instructions")
            out.append(f"  * hardware paths: {', '.join(keys) if keys else '-'}")
            out.append(f"  * most frequent: {top}")
        out.append("")
    path = os.path.join(ROOT, "profiles", f"{args.tag}_sass_histogram.md")
    open(path, "w").write("\n".join(out) + "\n")
    print(path)


if __name__ == "__main__":
    main()
