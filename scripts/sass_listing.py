"""profiles/sass_gemm_kernel.txt: per gemm_kernel instantiation, the counts of the SASS mnemonics that show the
tcgen05 / TMA / mbarrier / PDL machinery (cuobjdump -sass of the in-tree object)."""
import collections, os, re, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
obj = os.path.join(root, "summarizer_b200", "csrc", "smz_gemm.o")
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
out = ["# cuobjdump -sass summarizer_b200/csrc/smz_gemm.o | grep -E 'UTC|UTMA|LDTM|SYNCS|ELECT|ACQBULK|PREEXIT' — per kernel instantiation",
       "# (gemm_kernel<A_MN, B_MN, PAIR, EPI>: EPI 0 = plain, 1 = head, 2 = exp, 3 = plain with hi + lo output planes).",
       "# UTCHMMA = tcgen05.mma kind::f16, .2CTA = cta_group::2, UTMALDG = cp.async.bulk.tensor (TMA) load, LDTM = tcgen05.ld,",
       "# UTCBAR = tcgen05.commit (multicast to both CTAs of the pair), ACQBULK / PREEXIT = griddepcontrol.wait / .launch_dependents (PDL).", ""]
for m in re.finditer(r"Function : (\S+)\n(.*?)(?=\n\s*Function : |\Z)", txt, flags=re.S):
    name, body = m.group(1), m.group(2)
    t = re.search(r"gemm_kernelILb(\d)ELb(\d)ELb(\d)ELi(\d)", name)
    if not t:
        continue
    ops = re.findall(r"^\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", body, flags=re.M)
    cnt = collections.Counter(o for o in ops if re.match(r"UTC|UTMA|LDTM|SYNCS|ELECT|R2UR|ACQBULK|PREEXIT", o))
    out.append(f"gemm_kernel<A_MN={t.group(1)}, B_MN={t.group(2)}, PAIR={t.group(3)}, EPI={t.group(4)}>   ({len(ops)} instructions)")
    for k in sorted(cnt):
        out.append(f"    {k:<40} x{cnt[k]}")
    out.append("")
open(os.path.join(root, "profiles", "sass_gemm_kernel.txt"), "w").write("\n".join(out))
print(len(out), "lines")
