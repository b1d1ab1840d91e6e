import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np
from moonshine_b200 import scenes
from gpu_common import gpu_render, oracle_render, within_tol_fraction
sc = scenes.bathroom((160, 90), detail=0.03)
g, st, _ = gpu_render(sc, 1); r, c, _ = oracle_render(sc, 1)
d = np.abs(g[...,:3].astype(np.float64)-r[...,:3]); tol = 1e-4*np.maximum(1,np.abs(r[...,:3]))
bad = (d>tol).any(-1)
print('frac ok', within_tol_fraction(g,r), 'bad pixels', bad.sum(), 'of', bad.size, 'rays', st.rays_closest, c[0], st.rays_shadow, c[1])
rel = d/np.maximum(1,np.abs(r[...,:3]))
print('rel err percentiles', np.percentile(rel,[50,90,99,99.9,100]))
ys,xs = np.nonzero(bad); print(list(zip(xs[:20],ys[:20])))
for (x,y) in list(zip(xs[:8],ys[:8])): print(x,y,g[y,x,:3],r[y,x,:3])
for mc in [(0,0),(1,0),(0,1)]:
    sc.constants.env_samples_per_bounce, sc.constants.mesh_samples_per_bounce = mc
    g, st, _ = gpu_render(sc, 1); r, c, _ = oracle_render(sc, 1)
    print('env,mesh',mc,'frac', within_tol_fraction(g,r), st.rays_closest, c[0])
sc.constants.env_samples_per_bounce = sc.constants.mesh_samples_per_bounce = 1
sc.constants.max_bounces = 0
g, st, _ = gpu_render(sc, 1); r, c, _ = oracle_render(sc, 1)
print('max_bounces 0 frac', within_tol_fraction(g,r), st.rays_closest, c[0])
