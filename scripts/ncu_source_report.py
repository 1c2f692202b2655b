"""Summarises an `ncu --page source --csv` dump (SASS view): where the warp-stall samples and the executed
instructions of a kernel sit, bucketed along the instruction stream, with the dominant stall reason per bucket.

    python scripts/ncu_source_report.py <source.csv> [--bucket 128] [--top 25]
"""
import argparse
import csv
import collections


def load(path):
  rows = list(csv.reader(open(path)))
  hdr = rows[1]
  col = {h: i for i, h in enumerate(hdr)}
  stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
  ins = []
  for r in rows[2:]:
    if len(r) < len(hdr):
      continue
    ins.append(dict(sass=r[col['Source']].strip(), samples=int(r[col['# Samples']] or 0),
                    executed=int(r[col['Instructions Executed']] or 0),
                    stalls={h: int(r[col[h]] or 0) for h in stall_cols}))
  return ins


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('path')
  ap.add_argument('--bucket', type=int, default=128)
  ap.add_argument('--top', type=int, default=25)
  args = ap.parse_args()
  ins = load(args.path)
  total_s = sum(i['samples'] for i in ins); total_e = sum(i['executed'] for i in ins)
  print(f'{len(ins)} SASS instructions, {total_e} warp instructions executed, {total_s} stall samples')
  agg = collections.Counter()
  for i in ins:
    agg.update(i['stalls'])
  print('stall reasons:', ', '.join(f'{k[6:]} {100 * v / max(1, total_s):.1f}%' for k, v in agg.most_common(8)))
  print(f'\n{"bucket":>12s} {"exec %":>7s} {"samples %":>9s}  top stalls / marker instructions')
  for b in range(0, len(ins), args.bucket):
    chunk = ins[b:b + args.bucket]
    e = sum(i['executed'] for i in chunk); s = sum(i['samples'] for i in chunk)
    st = collections.Counter()
    for i in chunk:
      st.update(i['stalls'])
    marks = [i['sass'].split()[0] if not i['sass'].startswith('@') else i['sass'].split()[1] for i in chunk]
    marks = [m for m in marks if m.split('.')[0] in ('BAR', 'UBLKCP', 'SYNCS', 'EXIT', 'MUFU', 'LDG', 'STG', 'DFMA')]
    mc = collections.Counter(m.split('.')[0] for m in marks)
    print(f'{b:6d}-{b + len(chunk):<6d} {100 * e / max(1, total_e):7.2f} {100 * s / max(1, total_s):9.2f}  '
          + ' '.join(f'{k[6:]}:{100 * v / max(1, s):.0f}%' for k, v in st.most_common(3)) + '   ' + dict(mc).__repr__())
  print('\ntop instructions by samples:')
  order = sorted(range(len(ins)), key=lambda k: -ins[k]['samples'])[:args.top]
  for k in order:
    i = ins[k]
    top = max(i['stalls'].items(), key=lambda kv: kv[1])
    print(f'{k:6d} {100 * i["samples"] / max(1, total_s):5.2f}% exec {i["executed"]:9d} {top[0][6:]:12s} {i["sass"][:90]}')


if __name__ == '__main__':
  main()
