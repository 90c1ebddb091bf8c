import time, numpy as np, sys
sys.path.insert(0, '.')
t0=time.time()
from votca_b200.api import Context
ctx = Context(0); print('ctx', time.time()-t0)
rng = np.random.default_rng(0)
def T(label, f):
    t=time.time(); r=f(); ctx.sync(); print(f'{label}: {time.time()-t:.4f}s'); return r
A = rng.standard_normal((257,100)); B = rng.standard_normal((100,129))
for i in range(3):
    dA = T('upload', lambda: ctx.upload(A)); dB = ctx.upload(B); dC = ctx.upload(np.zeros((257,129)))
    for cfg in (0,1,2,-1):
        T(f'dgemm cfg{cfg}', lambda: ctx.dgemm('N','N',257,129,100,1.0,dA,257,dB,100,0.0,dC,257,cfg,0))
    T('download', lambda: ctx.download(dC,(257,129)))
    T('free', lambda: [ctx.free(p) for p in (dA,dB,dC)])
S = A.T@A + np.eye(100)
for i in range(2):
    T('sym_eig', lambda: ctx.sym_eig(S))
    T('inverse', lambda: ctx.inverse(S))
    T('gen_eig', lambda: ctx.gen_eig(S[:20,:20], S[:20,:20]+rng.standard_normal((20,20))))
# big gemm perf
for n in (2048, 4096, 8192):
    dA = ctx.upload(rng.standard_normal((n,n))); dB = ctx.upload(rng.standard_normal((n,n))); dC = ctx.malloc(n*n)
    for (ta,tb) in (('N','N'),('T','N'),('N','T'),('T','T')):
        ctx.dgemm(ta,tb,n,n,n,1.0,dA,n,dB,n,0.0,dC,n,0,1); ctx.sync()
        ctx.timer_start()
        for r in range(3): ctx.dgemm(ta,tb,n,n,n,1.0,dA,n,dB,n,0.0,dC,n,0,1)
        ms = ctx.timer_stop_ms()/3
        print(f'dgemm {ta}{tb} n={n}: {ms:.3f} ms  {2*n**3/ms/1e9:.2f} TFLOP/s')
    for p in (dA,dB,dC): ctx.free(p)
