import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'oracle'))
import numpy as np, gtn_oracle as O, grassmanntn_b200 as gtn
g = gtn.gauge2d
rng = np.random.RandomState(21)
a = O.random_dense((4,4,4,4),(1,1,-1,-1),dtype=complex,rng=rng)
A = gtn.dense(a.data, statistics=(1,1,-1,-1))
sw = lambda X: O.einsum("jikl->jilk", O.einsum("ijkl->jikl", X))
as_ = sw(a); As = g._swap_xy(A)
print('swap diff', np.abs(As.data.cpu().numpy()-as_.data).max())
for cut in (8, 16, None):
    for et in (False, True):
        ro = O.atrg2dy(as_, as_, cut, error_test=et); rg = g.atrg2dy(As, As, cut, error_test=et)
        print('atrg2dy(swapped)', cut, et, ro[1], rg[1])
As2 = gtn.dense(as_.data, statistics=(1,1,-1,-1))
rg = g.atrg2dy(As2, As2, 8); print('fresh dense swapped', rg[1])
ro = O.atrg2dx(a, a, 8); rg = g.atrg2dx(A, A, 8); print('atrg2dx', ro[1], rg[1])
