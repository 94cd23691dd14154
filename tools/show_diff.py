import numpy as np, sys, os
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from helpers import *
from oracle.oracle import f32_record_to_words
from dynenv_b200 import layout as L
name=sys.argv[1]; envs=[int(x) for x in sys.argv[2].split(',')]; fields=sys.argv[3].split(',')
d=np.load('gpurun_out/dump_%s.npz'%name); g=load_golden(name); cfg=golden_cfg(g)
outs,ow=oracle_step_batch(cfg,d['recs'],g['actions'],int(d['seed']))
gw=f32_record_to_words(d['state_after'])
w0=f32_record_to_words(d['recs'])
F=L.RC_FIELDS
np.set_printoptions(linewidth=200, precision=6, suppress=True)
def arb(w):
    n=int(w[F['nArb'][0]]); r=[]
    for k in range(n):
        key=int(w[F['arbKey'][0]+k]); con=int(w[F['arbCon'][0]+k])
        r.append((key&255,(key>>8)&255,'st%d'%((key>>16)&255),'age%d'%((key>>24)&255),'n%d'%(con&255),(con>>8)&255,(con>>16)&255,
                  round(float(w[F['arbJn0'][0]+k]),2),round(float(w[F['arbJt0'][0]+k]),2),round(float(w[F['arbJn1'][0]+k]),2),round(float(w[F['arbJt1'][0]+k]),2)))
    return r
for e in envs:
    print('env',e,'actions',g['actions'][e].tolist())
    for n in fields:
        if n=='arb':
            print(' before',arb(w0[e])); print(' gpu   ',arb(gw[e])); print(' oracle',arb(ow[e])); continue
        if n in F:
            o,c,k=F[n]
            print(n,'\n  gpu',gw[e,o:o+c],'\n  orc',ow[e,o:o+c])
        else:
            print(n,'\n gpu',d[n][e],'\n orc',outs[e][n])
