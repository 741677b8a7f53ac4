"""Soak run of the host model of the Fano queue protocol (tools/fano_queue_host_check.cpp, tests/test_fano_queue_host.py):
random shapes -- 1-9 contexts, 4-47 captures each, 200-1500 candidates per context, bursts up to 32, rings of 512-2048 entries,
pools of 1-23 workers with and without a per-SM cap, quick-mode candidates mixed in, random yields at every atomic and fence,
easy and hopeless symbol vectors -- for the given number of seconds.   python tools/fano_queue_soak.py 1500
(profiles/r2_fano_queue_model.txt: 698 runs, 0 failures)"""
import sys, ctypes as C, numpy as np, time, os, subprocess, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, ROOT)
from test_fano_host import vectors
CSRC=os.path.join(ROOT, 'rtlsdr_wsprd_b200', 'csrc')
src=open(CSRC+'/wspr_kernels.cu').read()
a=src.index("__device__ void fano_settle(ChainScratch *cs, int count) {"); b=src.index("// The pool: at most q->pool worker warps are alive at any time.")
code=src[a:b]
rh="            unsigned h = *(volatile unsigned *)&q->head0;\n"
libs=[]
for name, c in (("v", code), ("w", code.replace(rh, "            sched_yield();\n"+rh))):
    d=tempfile.mkdtemp(prefix="fqsoak_"); open(d+'/fano_queue_extracted.inc','w').write(c)
    out=d+'/lib.so'
    subprocess.run(["g++","-O2","-std=c++17","-pthread","-shared","-fPIC","-I/usr/local/cuda/include","-I"+CSRC,"-I"+d,"-o",out,os.path.join(ROOT, "tools", "fano_queue_host_check.cpp")],check=True)
    lib=C.CDLL(out); lib.fano_queue_sim.argtypes=[C.c_void_p]+[C.c_int]*10+[C.c_uint,C.c_uint]+[C.c_int]*3+[C.c_void_p]; libs.append(lib)
easy=np.ascontiguousarray(np.stack(vectors(64,7)))
hard=np.ascontiguousarray(np.stack([v for i,v in enumerate(vectors(256,11)) if i%8>=5][:64]))
rng=np.random.default_rng(12345)
fails=0; t0=time.time(); runs=0
while time.time()-t0 < float(sys.argv[1]):
    nctx=int(rng.integers(1,10)); ncap=int(rng.integers(4,48)); ncand=int(rng.integers(200,1500)); burst=int(rng.integers(1,33))
    ring=int(rng.integers(9,12)); pool=int(rng.integers(1,24)); per_sm=int(rng.choice([0,1,2,3])); nsm=int(rng.integers(1,12))
    if per_sm>0 and per_sm*nsm < 1: continue
    maxc=int(rng.choice([20,30,60,150])); seed=int(rng.integers(1<<30)); qe=int(rng.choice([0,0,3,5,9])); chaos=int(rng.choice([0,2,3,8]))
    vecs=easy if rng.random()<0.6 else hard
    lib=libs[int(rng.integers(2))]
    out=np.zeros(8,np.int64)
    rc=lib.fano_queue_sim(vecs.ctypes.data,64,nctx,ncap,ncand,burst,ring,pool,per_sm,nsm,60,maxc,seed,30000,qe,chaos,out.ctypes.data)
    runs+=1
    if rc!=0:
        fails+=1; print("FAIL rc",rc,dict(nctx=nctx,ncap=ncap,ncand=ncand,burst=burst,ring=ring,pool=pool,per_sm=per_sm,nsm=nsm,maxc=maxc,seed=seed,qe=qe,chaos=chaos,hard=vecs is hard),out[:5].tolist(),flush=True)
print("runs",runs,"failures",fails,"seconds",round(time.time()-t0),flush=True)
