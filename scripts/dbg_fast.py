import sys, numpy as np, torch
sys.path.insert(0, '.')
import optimobo_b200 as ob
from oracle import oracle as O
def zdt1(X):
    f1 = X[:, 0]; g = 1 + 9.0/(X.shape[1]-1)*X[:,1:].sum(1)
    return np.column_stack([f1, g*(1-np.sqrt(f1/g))])
for (n,d,m) in [(100,4,128),(128,10,1000),(256,10,4096),(512,12,4096),(1024,10,8192),(700,7,5000)]:
    rng=np.random.default_rng(0); X=rng.random((n,d)); Y=zdt1(X)
    for i,(ell,sf2) in enumerate([(0.7,1.0),(0.8,2.0)]):
        gp=ob.GPModel(X,Y[:,i],ell*np.ones(d),sf2,device='cuda:0')
        st=O.gp_fit_state(X,Y[:,i],ell*np.ones(d),sf2)
        Xc=rng.random((m,d))
        mu_o,var_o=O.gp_posterior(st,Xc)
        mu,var=ob.posterior([gp],Xc,precision='fast')
        torch.cuda.synchronize()
        mu=mu[0].cpu().numpy(); var=var[0].cpu().numpy()
        sd=np.sqrt(var); sdo=np.sqrt(var_o)
        print(f"n={n} d={d} m={m} gp{i}: mu abs err {np.abs(mu-mu_o).max():.3g}  sd rel err max {np.abs(sd-sdo).max()/1:.3g} / rel {(np.abs(sd-sdo)/sdo).max():.3g}", flush=True)
