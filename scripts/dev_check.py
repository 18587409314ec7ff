"""Developer check run on the GPU box: compares the CUDA path with the oracle step by step."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dvo_slam_b200 import synth
from dvo_slam_b200.engine import Engine, Config
from oracle import oracle_py as orc

def nan_eq(a, b):
    return np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(a[~np.isnan(a)], b[~np.isnan(b)])

eng = Engine(0)
p = synth.make_pair(0)
K = p['intrinsics']
arrs = {k: p[k].numpy() for k in ('I_ref','Z_ref','I_cur','Z_cur')}
L = 5
gref = eng.pyramid(arrs['I_ref'], arrs['Z_ref'], K, L); gcur = eng.pyramid(arrs['I_cur'], arrs['Z_cur'], K, L)
oref = orc.Pyramid(arrs['I_ref'], arrs['Z_ref'], K, L); ocur = orc.Pyramid(arrs['I_cur'], arrs['Z_cur'], K, L)
for l in range(L):
    g = gref.download(l); o = oref.planes(l)
    bad = np.isnan(o).any(axis=0)
    oz = o.copy(); oz[1][bad] = np.nan
    ok = all(nan_eq(g[c], oz[c]) for c in range(6))
    S, mask = gref.select(l); So, masko = orc.select(oref, l)
    print("level", l, "planes bitexact", ok, "S", S, So, "mask eq", np.array_equal(mask, masko))
mir = orc.mode('mirror')
T = np.linalg.inv(np.eye(4)); T = synth.se3_exp(p['xi']*0.9)
for l in [4, 2, 0]:
    n_g, img_g = eng.residual_image(gref, gcur, l, T)
    n_o, img_o = orc.residual_image(oref, ocur, l, T, mir)
    print("resid level", l, "n", n_g, n_o, "bitexact", nan_eq(img_g, img_o), "maxabs", np.nanmax(np.abs(img_g-img_o)) if n_g else None)
    for uw in (False, True):
        pp = np.array([[2000., -30.],[-30., 9000.]], dtype=np.float32)
        lg = eng.linearize(gref, gcur, l, T, uw, pp); lo = orc.linearize(oref, ocur, l, T, mir, uw, pp)
        print("  lin uw", uw, "n", lg['n'], lo['n'], "P rel", np.abs(lg['precision']-lo['precision']).max()/np.abs(lo['precision']).max(),
              "ll", lg['ll'], lo['ll'], "A rel", np.abs(lg['A']-lo['A']).max()/np.abs(lo['A']).max(), "b rel", np.abs(lg['b']-lo['b']).max()/np.abs(lo['b']).max())
cfg = Config(first_level=4, last_level=0, max_iterations_per_level=50, precision=1e-4)
ocfg = orc.config(first_level=4, last_level=0, max_iterations_per_level=50, precision=1e-4)
t=time.time(); rg = eng.match(gref, gcur, cfg, with_iterations=True); tg=time.time()-t
t=time.time(); rg = eng.match(gref, gcur, cfg, with_iterations=True); tg2=time.time()-t
for name in ('mirror','faithful'):
    t=time.time(); ro = orc.match(oref, ocur, ocfg, orc.mode(name)); to=time.time()-t
    d = synth.se3_log(np.linalg.inv(ro['T']) @ rg.transformation)
    print(name, "dT %.2e dR %.2e" % (np.abs(d[:3]).max(), np.abs(d[3:]).max()), "its gpu", [l['num_iterations'] for l in rg.levels], "orc", [l['num_iterations'] for l in ro['levels']],
          "tc", [l['termination'] for l in rg.levels], [l['termination'] for l in ro['levels']], "info ratio", rg.information[0,0]/ro['information'][0,0], "ll", rg.log_likelihood, ro['log_likelihood'], "t_orc %.3f t_gpu %.4f %.4f" % (to, tg, tg2))
gt = synth.se3_log(rg.transformation @ p['T_true'])
print("gpu vs truth", np.abs(gt[:3]).max(), np.abs(gt[3:]).max())
# batch
B = 16
pairs = [synth.make_pair(s) for s in range(B)]
Ir = np.stack([q['I_ref'].numpy() for q in pairs]); Zr = np.stack([q['Z_ref'].numpy() for q in pairs])
Ic = np.stack([q['I_cur'].numpy() for q in pairs]); Zc = np.stack([q['Z_cur'].numpy() for q in pairs])
refs = eng.pyramid_batch(Ir, Zr, K, L); curs = eng.pyramid_batch(Ic, Zc, K, L)
eng.profile_enable(True)
t=time.time(); res = eng.match_batch(refs, curs, cfg); tb=time.time()-t
print("batch", B, "time", tb, eng.profile_read())
for i, r in enumerate(res):
    o_r = orc.Pyramid(Ir[i], Zr[i], K, L); o_c = orc.Pyramid(Ic[i], Zc[i], K, L)
    rm = orc.match(o_r, o_c, ocfg, orc.mode('mirror')); rf = orc.match(o_r, o_c, ocfg, orc.mode('faithful'))
    dm = synth.se3_log(np.linalg.inv(rm['T']) @ r.transformation); df = synth.se3_log(np.linalg.inv(rf['T']) @ r.transformation)
    print(i, "vs mirror %.1e %.1e" % (np.abs(dm[:3]).max(), np.abs(dm[3:]).max()), "vs faithful %.1e %.1e" % (np.abs(df[:3]).max(), np.abs(df[3:]).max()),
          [l['num_iterations'] for l in r.levels], [l['num_iterations'] for l in rm['levels']], [l['num_iterations'] for l in rf['levels']])
