"""Per-kernel SASS listings of the Blackwell-specific kernels: `python tools/sass_listing.py` rewrites docs/sass/*.sass,
docs/sass/kernels.txt and docs/sass/opcode_census.md from the in-tree build (cuobjdump -sass of feddrift_b200/_C/fdb200_C.so).

Each listing keeps the instructions that prove the hardware path (tcgen05.mma = UTCHMMA, tcgen05.ld/st = LDTM/STTM, tcgen05.commit =
UTCBAR, TMA = UTMALDG/UTMASTG/UTMAREDG/UBLKCP, mbarrier = SYNCS, cluster barriers = UCGABAR_*, DSMEM = MAPA / ST.*cluster, NVLS =
LDGMC/multimem, peer stores) with two lines of context each, plus the opcode histogram of the whole kernel."""
import collections
import os
import re
import subprocess
import sys

SO = sys.argv[1] if len(sys.argv) > 1 else "feddrift_b200/_C/fdb200_C.so"
OUT = "docs/sass"
KEY = re.compile(r"\b(UTC[A-Z]*MMA|UTCBAR|UTCCP|LDTM|STTM|UTMALDG|UTMASTG|UTMAREDG|UBLKCP|UBLKRED|SYNCS|UCGABAR_ARV|UCGABAR_WAIT|MAPA|LDGMC|"
                 r"MULTIMEM|REDG|ATOMG|HMMA|MUFU\.TANH|CCTL|ERRBAR|MEMBAR\.[A-Z.]*SYS|LDGSTS)\b")
WANT = ["gemm_tn_kernel", "lstm2_fwd_kernel", "lstm2_bwd_kernel", "lstm_head_kernel", "lstm_small_grads_kernel", "conv_igemm_kernel", "conv_wgrad_kernel",
        "fed_round_small_kernel", "fedavg_reduce_apply_peer_kernel", "gossip_mix_peer_kernel", "cluster_aggregate_kernel", "gram_kernel",
        "modp_matmul_mont_kernel", "group_norm_bwd_kernel"]

sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
funcs, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = []
        continue
    if cur and re.match(r"\s*/\*[0-9a-f]{4}\*/", line):
        funcs[cur].append(re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", line).rstrip())


def demangle(names):
    try:
        out = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.splitlines()
        return out if len(out) == len(names) else names
    except OSError:
        return names


names = list(funcs)
pretty = dict(zip(names, demangle(names)))
os.makedirs(OUT, exist_ok=True)
with open(os.path.join(OUT, "kernels.txt"), "w") as fh:
    for n in names:
        fh.write(f"{pretty[n]}    [{len(funcs[n])} SASS instructions]\n")
seen = set()
for n in names:
    base = next((w for w in WANT if w in n), None)
    if base is None:
        continue
    tag = re.sub(r"[^A-Za-z0-9_]+", "_", pretty[n].split("(")[0].replace("void ", "").replace("fdb::", ""))[:80]
    if base in ("fed_round_small_kernel", "gram_kernel", "conv_igemm_kernel", "conv_wgrad_kernel", "gemm_tn_kernel") and base in seen:
        continue          # one representative instantiation per template
    seen.add(base)
    lines = funcs[n]
    hist = collections.Counter(re.sub(r"^\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?", "", ln).split(" ")[0].split(";")[0] for ln in lines)
    keep = set()
    for i, ln in enumerate(lines):
        if KEY.search(ln):
            keep.update(range(max(0, i - 2), min(len(lines), i + 3)))
    with open(os.path.join(OUT, f"{tag}.sass"), "w") as fh:
        fh.write(f"// {pretty[n]}\n// {len(lines)} SASS instructions (sm_100a); excerpt: Blackwell-specific instructions with 2 lines of context\n")
        fh.write("// opcode histogram: " + ", ".join(f"{k}×{v}" for k, v in hist.most_common(40)) + "\n\n")
        last = -2
        for i in sorted(keep):
            if i != last + 1:
                fh.write("        ...\n")
            fh.write(lines[i] + "\n")
            last = i
print("wrote", len(seen), "listings to", OUT)
