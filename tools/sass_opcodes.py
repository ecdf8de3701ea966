"""Per-kernel SASS opcode histogram of libduoformer_sm100.so (cuobjdump -sass; runs without a GPU): the evidence that
the hot kernels are tcgen05 / TMEM / TMA code.  Writes profiles/sass_opcodes.txt.
  UTCHMMA[.2CTA]  tcgen05.mma (cta_group::1 / ::2)      LDTM            tcgen05.ld (TMEM -> registers)
  UTMALDG         TMA load (cp.async.bulk.tensor)       UTMASTG         TMA store
  UTMAREDG        TMA reduce-add                        UTCBAR          tcgen05.commit -> mbarrier (MULTICAST: both CTAs)
  HMMA            legacy mma.sync tensor path           MUFU / FFMA2    special-function / packed fp32 math"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "duoformer_tcga_b200", "libduoformer_sm100.so")
KEYS = ["UTCHMMA.2CTA", "UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTCBAR", "HMMA", "LDSM", "MUFU", "FFMA2", "FFMA",
        "SYNCS", "STS", "LDS", "LDG", "STG"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    funcs = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*(?:\.[A-Z0-9_]+)*)", line)
        if m and cur:
            op = m.group(1)
            funcs[cur]["_total"] += 1
            for k in KEYS:
                if op == k or op.startswith(k + "."):
                    if k == "UTCHMMA" and ".2CTA" in op:
                        continue
                    if k == "FFMA" and op.startswith("FFMA2"):
                        continue
                    funcs[cur][k] += 1
            if ".2CTA" in op and op.startswith("UTCHMMA"):
                funcs[cur]["UTCHMMA.2CTA"] += 0  # counted by its own key above
            if op.startswith("UTCBAR") and "MULTICAST" in op:
                funcs[cur]["UTCBAR.MULTICAST"] += 1
            if op.startswith("UTMALDG") and ".2CTA" in op:
                funcs[cur]["UTMALDG.2CTA"] += 1
    names = subprocess.run(["c++filt"], input="\n".join(funcs), capture_output=True, text=True).stdout.splitlines()
    out = [__doc__.strip(), "", f"library: {os.path.relpath(LIB, ROOT)} ({os.path.getsize(LIB)} bytes), {len(funcs)} kernels", ""]
    total = collections.Counter()
    cols = KEYS[:10] + ["UTCBAR.MULTICAST", "UTMALDG.2CTA"]
    out.append("instr  " + " ".join(c.rjust(12) for c in cols) + "  kernel")
    for (mangled, c), name in zip(funcs.items(), names):
        name = re.sub(r"\(anonymous namespace\)::", "", name)
        name = re.sub(r"\(CUtensorMap_st.*", "(...)", name)
        out.append(f"{c['_total']:6d} " + " ".join(str(c[k]).rjust(12) for k in cols) + "  " + name[:110])
        total.update(c)
    out.append("")
    out.append("totals: " + ", ".join(f"{k} {total[k]}" for k in cols))
    path = os.path.join(ROOT, "profiles", "sass_opcodes.txt")
    open(path, "w").write("\n".join(out) + "\n")
    print("\n".join(out[-3:]))
    print("wrote", path)


if __name__ == "__main__":
    sys.exit(main())
