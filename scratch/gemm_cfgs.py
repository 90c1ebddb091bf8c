import sys, numpy as np
sys.path.insert(0, '.')
from votca_b200.api import Context
ctx = Context(0)
def bench(ta, tb, m, n, k, cfg, splitk=1, reps=3):
    A = ctx.malloc(m*k); B = ctx.malloc(k*n); C = ctx.malloc(m*n)
    lda = k if ta=='T' else m; ldb = n if tb=='T' else k
    ctx.dgemm(ta, tb, m, n, k, 1.0, A, lda, B, ldb, 0.0, C, m, cfg, splitk); ctx.sync()
    ctx.timer_start()
    for r in range(reps): ctx.dgemm(ta, tb, m, n, k, 1.0, A, lda, B, ldb, 0.0, C, m, cfg, splitk)
    ms = ctx.timer_stop_ms()/reps
    for p in (A,B,C): ctx.free(p)
    return 2*m*n*k/ms/1e9
for cfg in (0,3,4,5):
    print('cfg', cfg, 'TN 4096^3: %.2f' % bench('T','N',4096,4096,4096,cfg), ' NN: %.2f' % bench('N','N',4096,4096,4096,cfg),
          ' TN 4320x50832x287: %.2f' % bench('T','N',4320,50832,287,cfg), ' TN 144x2880xK=65536: %.2f' % bench('T','N',144,2880,65536,cfg,4),
          ' TN 287x2880x65536: %.2f' % bench('T','N',287,2880,65536,cfg,4), flush=True)
