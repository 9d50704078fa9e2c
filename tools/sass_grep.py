"""profiles/sass_grep.txt: per kernel of libfgk_b200.so, how many SASS instructions of the kinds
that prove the design (TMA bulk copy + mbarrier, 128-bit / streaming loads, atomics, system-scope
fences and flag accesses of the peer kernels, ballot / match / popc / redux of the enumerators)."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "flow_guided_krylov_b200", "csrc", "libfgk_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
pat = re.compile(r"\b(UBLKCP\S*|SYNCS\S*|LDG\.E\.(?:NA\.)?128\S*|LDG\.E\.(?:NA\.)?64\S*|LDG\.E\.ENL2\S*|STG\.E\.128\S*|ATOMG\S*|ATOM\.\S*|RED\.\S*|"
                 r"MEMBAR\S*|LD\.E\.\S*SYS\S*|ST\.E\.\S*SYS\S*|LDG\.E\.\S*SYS\S*|STG\.E\.\S*SYS\S*|MATCH\S*|VOTE\S*|REDUX\S*|POPC|FLO\S*|DFMA|CCTL\S*)")
fn, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = subprocess.run(["c++filt", "-p", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        counts[fn] = collections.Counter()
        continue
    if fn:
        for tok in pat.findall(line):
            counts[fn][tok.rstrip(",;")] += 1
print("# cuobjdump -sass flow_guided_krylov_b200/csrc/libfgk_b200.so (sm_100a), instruction counts per kernel")
print("# UBLKCP = cp.async.bulk (TMA bulk copy), SYNCS = mbarrier; LDG.E.NA.128 = ld.global.nc.L1::no_allocate.v4 / v2.f64")
print("# (the streaming matrix loads); .SYS = system-scope (peer memory) accesses")
for fn, c in counts.items():
    if c:
        print(f"{fn}\n    " + "  ".join(f"{k} x{v}" for k, v in sorted(c.items())))
