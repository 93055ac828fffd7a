"""Blackwell-native evidence: per kernel of libbq_b200.so, the count of the SASS mnemonics that prove tcgen05 / TMEM / TMA use
(B200_PROFILING.md "What proves a Blackwell-native kernel").  Usage: python tools/sass_listing.py > profiles/r02_sass_mnemonics.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "llm_mixed_q_b200", "libbq_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
pat = re.compile(r"\b(UTC[A-Z]*MMA(?:\.2CTA)?|UTCBAR(?:\.2CTA)?(?:\.MULTICAST)?|LDTM|STTM|UTMALDG(?:\.\dD)?|UTMASTG(?:\.\dD)?|UBLKCP|UTMAPF|SYNCS|HMMA|MUFU\.EX2|HFMA2\.BF16_V2|FMNMX3|LDGSTS)\b")
cur, counts, sizes = None, collections.OrderedDict(), {}
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur)
        counts[cur] = collections.Counter(); sizes[cur] = 0
        continue
    if cur and re.match(r"\s*/\*[0-9a-f]{4,}\*/", line):
        sizes[cur] += 1
        for k in pat.findall(line):
            counts[cur][k] += 1
print("# cuobjdump -sass llm_mixed_q_b200/libbq_b200.so — mnemonic counts per kernel (sm_100a); instructions = SASS lines")
for k, c in counts.items():
    if sizes[k] < 50 and not c:
        continue
    print(f"{k}\n    instructions {sizes[k]}  " + "  ".join(f"{n} {v}" for n, v in sorted(c.items())))
