"""SASS evidence: per-kernel instruction counts by mnemonic from `cuobjdump -sass recstudio_b200/librsb200.so`
(tcgen05 = UTCHMMA / UTCBAR / LDTM, bulk-copy TMA = UBLKCP, 16-byte vector loads / stores, shared / global atomics ...).
    python tools/sass_summary.py > profiles/r02_sass_summary.txt        (CPU only: needs the CUDA toolkit, no GPU)"""
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(REPO, "recstudio_b200", "librsb200.so")
PAT = {"UTCHMMA": r"UTCHMMA", "UTCBAR": r"UTCBAR", "LDTM": r"LDTM", "UBLKCP": r"UBLKCP", "UTMALDG": r"UTMALDG",
       "LDG.128": r"LDG\.E\.[A-Z.]*128", "STG.128": r"STG\.E\.[A-Z.]*128", "LDS": r"\bLDS", "STS": r"\bSTS",
       "ATOMS": r"ATOMS", "ATOMG": r"ATOMG", "RED": r"\bRED\.", "SHFL": r"SHFL", "VOTE": r"VOTE", "FFMA": r"FFMA",
       "IMAD": r"IMAD", "MUFU": r"MUFU", "BAR": r"BAR\.SYNC"}
KEEP = re.compile(r"pair_fwd|bin_|draw_bin|shard_|scatter_kernel|attn_|score_gmax|topk_|fullsoftmax|uniform_|popular_|masked_|"
                  r"kmeans|radix|segment_|rows_update|count_kernel|scan_|resolve|gather_rows|score_ids")


def main():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    funcs = re.split(r"\n\s*Function : ", txt)[1:]
    names = [f.split("\n", 1)[0].strip() for f in funcs]
    dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    rows = []
    for f, d in zip(funcs, dem):
        d = re.sub(r"\((rsb::|const|unsigned|long|int|float|void|[A-Za-z_]+ \*).*$", "", d).replace("void ", "")[:100]
        if not KEEP.search(d):
            continue
        n = len(re.findall(r"/\*[0-9a-f]{4}\*/", f))
        rows.append((d, n, [len(re.findall(v, f)) for v in PAT.values()]))
    print("# SASS summary of recstudio_b200/librsb200.so (cuobjdump -sass, sm_100a only): static instruction counts per kernel")
    print("# tcgen05: UTCHMMA (MMA) / UTCBAR (commit) / LDTM (tcgen05.ld); bulk-copy TMA: UBLKCP (cp.async.bulk); UTMALDG = tensor-map TMA (unused)")
    print("kernel | instrs | " + " | ".join(PAT))
    for d, n, c in sorted(rows):
        print("%s | %d | %s" % (d, n, " | ".join(str(x) for x in c)))


if __name__ == "__main__":
    sys.exit(main())
