"""Count SASS instructions per source line / per inlined function for one kernel (from `nvdisasm -g -c`).
usage: python scripts/sass_lines.py file.asm kernel_substr [top]"""
import re, sys, collections
asm, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
cur_k, cur = None, None
by_line, by_file = collections.Counter(), collections.Counter()
inl = collections.Counter()
total = 0
for l in open(asm, errors="replace"):
    m = re.match(r"\s*\.section\s+\.text\.(\S+),", l)
    if m: cur_k = m.group(1); continue
    if cur_k is None or kern not in cur_k: continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        m2 = re.search(r'inlined at "([^"]+)", line (\d+)', m.group(3))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
        total += 1
        by_line[cur] += 1
        by_file[cur[0] if cur else None] += 1
print("kernel", kern, "instructions", total, "=", total * 16 // 1024, "KB")
for f, c in by_file.most_common(12): print("%7d  %s" % (c, f))
print()
for (k, c) in by_line.most_common(top): print("%7d  %s:%s" % (c, k[0], k[1]) if k else "%7d  ?" % c)
