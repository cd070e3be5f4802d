import importlib, sys, numpy as np, torch
sys.path.insert(0,'.'); sys.path.insert(0,'oracle')
import tone_oracle as orc
tb=importlib.import_module('t-one_b200')
w=tb.weights.init_weights(0)
C,B,D,n=2400,1024,32,3
eng=tb.Engine(w,chunk_samples=C,max_slots=B,max_batch=B)
W=orc.to_torch(w)
distinct=tb.synth.telephony_pcm(D,C*n,seed=500+B)
idx=(np.arange(B)*7)%D
pcm=np.ascontiguousarray(distinct[idx]); slots=eng.alloc_slots(B)
st=orc.zero_state(D); outs=[]; refs=[]
for i in range(n):
    lp,_=eng.step(slots,pcm[:,i*C:(i+1)*C]); outs.append(lp.copy())
    r,st=orc.step(W,torch.from_numpy(distinct[:,i*C:(i+1)*C].astype(np.int32)),st); refs.append(r.numpy())
lp=np.stack(outs); ref=np.stack(refs)[:,idx]
err=np.abs(lp-ref)
print("max err",err.max(), "at ref logprob", ref.flat[err.argmax()])
for lo in (-2,-4,-6,-8,-10,-14,-30):
    m=ref>lo
    print(f"ref>{lo}: n={m.sum()} max err {err[m].max():.4f} p99.99 {np.percentile(err[m],99.99):.4f}")
# per distinct signal the spread across batch positions
print("count > 0.05:", (err>0.05).sum(), "of", err.size)
