import sys, json, time
sys.path.insert(0, 'tools'); sys.path.insert(0, '.')
import learning_curve as lc
for T, total in ((25, 2_000_000), (50, 2_000_000)):
    lc.TOTAL, lc.EVAL_EVERY, lc.T = total, 200_000, T
    t0 = time.time()
    curves = [lc.gpu_run(s, 64) for s in range(3)]
    print(json.dumps({"time_limit": T, "seconds": round(time.time() - t0, 1), "curves": [[(s, round(r, 3)) for s, r in c] for c in curves]}))
