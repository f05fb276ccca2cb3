"""Run the packaged tracker with the settings of the reference's published tables and print, per row, the largest relative
difference over the float columns and whether the integer columns agree (tests/golden/tracker_tables.json)."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import test_tracker_table as T

for table, bins_max in (("T25_sigma3", 512), ("adaptive", 512)):
    trk, args = T._args(table, bins_max)
    t0 = time.time()
    rows, reason = trk.run(args)
    print(f"{table}: {len(rows)} rows in {time.time() - t0:.1f} s (reference runtime_sec: "
          f"{sum(float(r['runtime_sec']) for r in T.TABLES[table]['rows']):.0f} s)")
    for got, ref in zip(rows, T.TABLES[table]["rows"]):
        ints = all(int(got[c]) == int(ref[c]) for c in T.INT_COLS) and all(got[c] == ref[c] for c in T.STR_COLS)
        worst = max((abs(float(got[c]) - float(ref[c])) / abs(float(ref[c])) if float(ref[c]) else abs(float(got[c])), c) for c in T.FLOAT_COLS)
        print(f"  bins={got['bins']}: int/str columns equal={ints}, worst float column {worst[1]} rel diff {worst[0]:.2e}, "
              f"kl_initial {got['kl_initial']!r} vs {ref['kl_initial']}, T_n {got['T_n']} vs {ref['T_n']}")
    args.out_prefix = str(ROOT / "gpurun_out" / f"tracker_{table}_b200")
    trk.write_outputs(args, rows, reason)
