"""profiles/rNN_sass_opcodes.txt: per-kernel counts of the SASS opcodes that prove the Blackwell paths
(UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG / UTMAREDG = TMA load / store / reduce,
UTCBAR = tcgen05.commit, SYNCS = mbarrier, UCGABAR = cluster barrier, LDS/STS .. for the DSMEM traffic) from
`cuobjdump -sass` of the shipped library."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "realtime-codec-agent_b200", "libmagicodec_b200.so")
OPS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTMAPF", "UTCBAR", "SYNCS", "UCGABAR_ARV",
       "UCGABAR_WAIT", "LD.E.128.STRONG", "HMMA", "MUFU.TANH", "MUFU.EX2", "FMNMX3"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = kernels.setdefault(re.sub(r"\(.*", "", name).replace("void ", "").replace("mc::", ""), collections.Counter())
            continue
        if cur is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            cur["total"] += 1
            for key in OPS:
                if op == key or op.startswith(key + ".") or (key in ("UTCHMMA.2CTA",) and key in op):
                    cur[key] += 1
    cols = [k for k in OPS if any(c[k] for c in kernels.values())]
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}   (cubin architectures: {', '.join(arch)})")
    print(f"{'kernel':78s} {'instrs':>7s} " + " ".join(f"{c[:11]:>11s}" for c in cols))
    for name, c in kernels.items():
        print(f"{name[:78]:78s} {c['total']:7d} " + " ".join(f"{c[k]:11d}" for k in cols))


if __name__ == "__main__":
    sys.exit(main())
