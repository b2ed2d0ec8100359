"""Developer script: run the CUDA path and the oracle side by side and print the differences."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mulls_b200 import synth, abi
from mulls_b200.registration import Context
from oracle import oracle

def compare(name, pair, ctx, reps=3):
    t = time.time()
    ro, to = oracle.icp_run(pair["tgt"], pair["src"], pair["params"], pair["init_guess"], threads=8)
    t_or = time.time() - t
    res, tr = ctx.run_batch([pair], want_trace=True)
    rg, tg = res[0], tr[0]
    print(f"== {name}: oracle {t_or*1e3:.1f} ms; code o/g {ro['code']}/{rg['code']} iters {ro['iters']}/{rg['iters']}")
    print("   n_corr oracle", ro["n_corr"], "gpu", rg["n_corr"])
    print("   n_src  oracle", ro["n_src"], "gpu", rg["n_src"])
    n = min(to["n_iter"], tg["n_iter"])
    for i in range(n):
        same = (to["n_corr"][i] == tg["n_corr"][i]).all() and (to["n_src"][i] == tg["n_src"][i]).all()
        da = np.abs(to["atpa"][i] - tg["atpa"][i]).max() / max(1e-30, np.abs(to["atpa"][i]).max())
        dx = np.abs(to["x"][i] - tg["x"][i]).max()
        print(f"   it{i}: counts_equal={same} rel|dATPA|={da:.2e} |dx|={dx:.2e}", "" if same else f"o={to['n_corr'][i]} {to['n_src'][i]} g={tg['n_corr'][i]} {tg['n_src'][i]}")
    dt, dr = synth.pose_error(rg["T"], ro["T"])
    print(f"   pose diff gpu-vs-oracle: {dt:.3e} m {dr:.3e} rad; sigma {ro['sigma']:.6g}/{rg['sigma']:.6g} conf {ro['confidence']:.6g}/{rg['confidence']:.6g}")
    print("   info rel diff", np.abs(ro["info"] - rg["info"]).max() / max(1e-30, np.abs(ro["info"]).max()))
    for _ in range(reps):
        ctx.upload([pair]); ctx.run_resident()
        print("   stats", ctx.stats())

if __name__ == "__main__":
    which = sys.argv[1:] or ["small", "c2"]
    ctx = Context(0, 4, 700000, 700000)
    for w in which:
        t = time.time(); pair = synth.make_pair(1000, w); print(w, "generated in", time.time() - t, [len(x) for x in pair["src"]])
        compare(w, pair, ctx)
