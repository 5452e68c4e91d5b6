import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'oracle'))
import numpy as np, gtn_oracle as O, grassmanntn_b200 as gtn
rng = np.random.RandomState(21)
a = O.random_dense((4,4,4,4),(1,1,-1,-1),dtype=complex,rng=rng)
A = gtn.dense(a.data, statistics=(1,1,-1,-1))
def cmp(name, g, r):
    gd = g.data.cpu().numpy(); rd = r.data
    print(name, gd.shape, rd.shape, g.statistics, r.statistics, np.abs(gd-rd).max() if gd.shape==rd.shape else 'SHAPE')
def both(sub, gs, rs):
    return gtn.einsum(sub,*gs), O.einsum(sub,*rs)
g1, r1 = both('ijkl->jikl',[A],[a]); cmp('sw1', g1, r1)
g1, r1 = both('jikl->jilk',[g1],[r1]); cmp('sw2', g1, r1)
T1g, T1r = both('ijkl->lijk',[g1],[r1]); cmp('T1', T1g, T1r)
Ug,Sg,Vg = T1g.svd('li|jk',8); Ur,Sr,Vr = O.svd(T1r,'li|jk',8)
print('S', np.diag(Sg.data.cpu().numpy()).real, np.diag(Sr.data).real)
recg, recr = both('lia,ab,bjk->lijk',[Ug,Sg,Vg],[Ur,Sr,Vr]); cmp('rec', recg, recr)
Bg, Br = both('lia,ab->lib',[Ug,Sg],[Ur,Sr])
Cg, Cr = both('ab,bjk->ajk',[Sg,Vg],[Sr,Vr])
Mg, Mr = both('ajk,jib->aibk',[Cg,Bg],[Cr,Br]); cmp('M', Mg, Mr)
U2g,S2g,V2g = Mg.svd('ai|bk',8); U2r,S2r,V2r = O.svd(Mr,'ai|bk',8)
print('S2', np.diag(S2g.data.cpu().numpy()).real, np.diag(S2r.data).real)
rec2g, rec2r = both('aix,xy,ybk->aibk',[U2g,S2g,V2g],[U2r,S2r,V2r]); cmp('rec2', rec2g, rec2r)
print('---- chain invariants')
def n2(g, r, name): print(name, g.norm, r.norm)
Dg, Dr = Ug, Ur; Ag, Ar = Vg, Vr
sqg, sqr = gtn.sqrt(S2g), O.sqrt(S2r); n2(sqg, sqr, 'sqrtS')
Yg, Yr = both('abx,xc->abc',[U2g,sqg],[U2r,sqr]); n2(Yg,Yr,'Y')
Xg, Xr = both('ax,xbc->abc',[sqg,V2g],[sqr,V2r]); n2(Xg,Xr,'X')
Q1g,Q1r = both('iax,xbj->ijab',[Dg,Yg],[Dr,Yr]); n2(Q1g,Q1r,'Q1')
Q2g,Q2r = both('kya,ylb->abkl',[Xg,Ag],[Xr,Ar]); n2(Q2g,Q2r,'Q2')
Qg,Qr = both('ijab,abkl->ijkl',[Q1g,Q2g],[Q1r,Q2r]); n2(Qg,Qr,'Q'); cmp('Q', Qg, Qr)
U3g,S3g,V3g = Qg.svd('ij|kl',8); U3r,S3r,V3r = O.svd(Qr,'ij|kl',8)
print('S3', np.diag(S3g.data.cpu().numpy()).real, np.diag(S3r.data).real)
