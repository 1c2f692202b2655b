"""Counts the SASS mnemonics that identify the hardware paths each kernel uses (TMA bulk copies = UBLKCP +
SYNCS mbarrier ops, tcgen05 tensor cores = UTCHMMA with TMEM loads LDTM and commits UTCBAR, TMA tensor loads / stores =
UTMALDG / UTMASTG, fp64 tensor cores = DMMA, 128-bit loads, shuffles, spills = STL / LDL) in the built
libble_b200.so.  CPU only:   python scripts/sass_mnemonics.py > profiles/r01c_sass_mnemonics.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'balloon_learning_environment_b200', 'libble_b200.so')
KEYS = ['UTCHMMA', 'LDTM', 'UTCBAR', 'UTMASTG', 'UBLKCP', 'UTMALDG', 'SYNCS', 'DMMA', 'HMMA', 'STG.E.EF.128', 'LDG.E.128', 'LDG.E.64', 'SHFL', 'DFMA', 'FFMA', 'MUFU', 'BAR.SYNC', 'STL', 'LDL']


def main():
  sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
  op = re.compile(r'^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)')
  counts, cur = collections.OrderedDict(), None
  for line in sass.splitlines():
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
      cur = m.group(1); counts[cur] = collections.Counter(); continue
    m = op.match(line)
    if m and cur:
      counts[cur][m.group(1)] += 1
  names = subprocess.run(['c++filt'], input='\n'.join(counts), capture_output=True, text=True).stdout.splitlines()
  print('%-40s %7s ' % ('kernel (sm_100a SASS)', 'instrs') + ' '.join('%9s' % k for k in KEYS))
  for (fn, c), name in zip(counts.items(), names):
    short = re.sub(r'^void ', '', re.sub(r'\(.*', '', name.replace('(anonymous namespace)::', '')))
    if not re.match(r'ble::k_', short):
      continue
    row = [sum(v for o, v in c.items() if o.startswith(k)) for k in KEYS]
    print('%-40s %7d ' % (short[:40], sum(c.values())) + ' '.join('%9d' % v for v in row))


if __name__ == '__main__':
  main()
