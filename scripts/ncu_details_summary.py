"""`ncu -i X.ncu-rep --page details --csv` -> a short text summary per launch (the sections the roofline claims rest on).
usage: ncu_details_summary.py details.csv > summary.txt"""
import collections
import csv
import sys

KEEP = ["Duration", "SM Frequency", "DRAM Frequency", "Memory Throughput", "DRAM Throughput", "L2 Cache Throughput",
        "L1/TEX Cache Throughput", "Compute (SM) Throughput", "Executed Ipc Active", "Issue Slots Busy", "SM Busy", "Mem Busy",
        "Max Bandwidth", "L1/TEX Hit Rate", "L2 Hit Rate", "Mem Pipes Busy", "No Eligible", "Eligible Warps Per Scheduler",
        "Warp Cycles Per Issued Instruction", "Avg. Active Threads Per Warp", "Registers Per Thread", "Block Size", "Grid Size",
        "Dynamic Shared Memory Per Block", "Static Shared Memory Per Block", "Waves Per SM", "Theoretical Occupancy",
        "Achieved Occupancy", "Block Limit Registers", "Block Limit Shared Mem"]
rows = list(csv.DictReader(open(sys.argv[1])))
by = collections.OrderedDict()
rules = collections.OrderedDict()
for r in rows:
    key = (r["ID"], r["Kernel Name"])
    if r["Metric Name"]:
        by.setdefault(key, collections.OrderedDict())[r["Metric Name"]] = (r["Metric Value"], r["Metric Unit"])
    if r["Rule Name"] in ("SOLBottleneck", "CPIStall", "HighPipeUtilization") and r["Rule Description"]:
        rules.setdefault(key, []).append(f'{r["Rule Name"]}: {r["Rule Description"][:400]}')
for key, m in by.items():
    print(f"== launch {key[0]}: {key[1][:150]}")
    for k in KEEP:
        if k in m:
            print(f"   {k:38s} {m[k][0]:>14s} {m[k][1]}")
    for t in rules.get(key, [])[:3]:
        print("   rule " + t)
    print()
