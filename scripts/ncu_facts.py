"""Turns ncu metric captures of `bench.py` into profiles/ncu_facts.json, the file bench.py reads its MEASURED
per-launch constants from (DRAM bytes of a k_wind_gather launch, executed warp instructions of a step-kernel launch).

    ncu --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,gpu__time_duration.sum \
        --clock-control none -k regex:'k_wind_gather|k_step_warp|k_step_roles' -c 40 --log-file <csv> python bench.py ...
    python scripts/ncu_facts.py <csv> [<csv> ...] --round r02

Nothing here is timed: numbers measured under ncu never become bench values; they are per-launch COUNTS.
"""
import argparse
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def launches(path):
  """-> list of dict(kernel, grid, block, metrics{name: value}) in launch order."""
  rows = {}
  with open(path, newline='') as f:
    lines = [l for l in f if l.startswith('"')]
  for r in csv.DictReader(lines):
    k = int(r['ID'])
    e = rows.setdefault(k, dict(kernel=r['Kernel Name'], grid=r['Grid Size'], block=r['Block Size'], metrics={}))
    try:
      v = float(r['Metric Value'].replace(',', ''))
    except ValueError:
      continue
    unit = r['Metric Unit']
    scale = {'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'byte': 1.0}.get(unit, 1.0)
    e['metrics'][r['Metric Name']] = v * scale
  return [rows[k] for k in sorted(rows)]


def dims(s):
  return [int(x) for x in re.findall(r'\d+', s)]


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('csv', nargs='+')
  ap.add_argument('--round', default='r02')
  ap.add_argument('--layout', default='x64')
  ap.add_argument('--num-envs', type=int, default=65536)
  ap.add_argument('--lookups', type=int, default=16777216)
  a = ap.parse_args()
  facts = {'round': a.round, 'gather_traffic': [], 'step_issue': None, 'launches': []}
  for path in a.csv:
    src = os.path.relpath(os.path.abspath(path), ROOT)
    for l in launches(path):
      m = l['metrics']
      name = re.sub(r'\(.*', '', l['kernel'])
      row = dict(kernel=name, grid=l['grid'], block=l['block'], source=src,
                 dram_bytes=m.get('dram__bytes_read.sum', 0.0) + m.get('dram__bytes_write.sum', 0.0),
                 warp_instructions=m.get('smsp__inst_executed.sum'), duration_ns=m.get('gpu__time_duration.sum'))
      facts['launches'].append(row)
  gather = [r for r in facts['launches'] if 'k_wind_gather' in r['kernel']]
  if gather:
    # bench.py launches the grouped-by-field order 13 times (3 warm-up + 10 timed), then the random-field order 12 times
    big = [r for r in gather if r['grid'] == gather[0]['grid']]
    for order, rows in (('grouped', big[:13]), ('random', big[13:25])):
      vals = sorted(r['dram_bytes'] for r in rows)
      if vals:
        facts['gather_traffic'].append(dict(layout=a.layout, lookups=a.lookups, order=order, grid=big[0]['grid'],
                                            dram_bytes=vals[len(vals) // 2], dram_bytes_min=vals[0], dram_bytes_max=vals[-1],
                                            launches=len(vals), source=big[0]['source']))
  step = [r for r in facts['launches'] if 'k_step' in r['kernel'] and r['warp_instructions']]
  if step:
    full = [r for r in step if dims(r['grid'])[0] * dims(r['block'])[0] >= a.num_envs]
    pick = full or step
    wi = sorted(r['warp_instructions'] for r in pick)[len(pick) // 2]
    db = sorted(r['dram_bytes'] for r in pick)[len(pick) // 2]
    facts['step_issue'] = dict(kernel=pick[0]['kernel'], grid=pick[0]['grid'], block=pick[0]['block'], num_envs=a.num_envs,
                               warp_instructions_per_launch=wi, warp_instructions_per_balloon=wi / a.num_envs,
                               thread_instructions_per_balloon=32.0 * wi / a.num_envs,
                               dram_bytes_per_launch=db, dram_bytes_per_balloon=db / a.num_envs,
                               launches=len(pick), source=pick[0]['source'])
  facts['launches'] = facts['launches'][:200]
  out = os.path.join(ROOT, 'profiles', 'ncu_facts.json')
  with open(out, 'w') as f:
    json.dump(facts, f, indent=1)
  print(json.dumps({k: facts[k] for k in ('gather_traffic', 'step_issue')}, indent=1))


if __name__ == '__main__':
  sys.exit(main())
