# validation of Tc3xCfg::SPLIT on the GPU: correctness per shape, then the GPU test-suite, then the size sweep per split mode
for sp in 2 1; do
export B200MM_TC3X_SPLIT=$sp
echo "== correctness SPLIT=$sp"
timeout 120 python -c "
import numpy as np, wgpu_mm_b200 as w, oracle
ctx=w.Context(0)
for (M,N,K) in [(256,512,384),(128,256,256),(1024,1024,1024),(4096,4096,512),(4096,4096,4096),(300,520,260)]:
  for t0 in (512,513):
    A=oracle.generate_weight_data(1,M,K); B=oracle.generate_weight_data(2,K,N)
    dA,dB=ctx.buffer_from(A),ctx.buffer_from(B); dC=ctx.buffer(M*N*4)
    k=ctx.kernel(w.KernelId.SGEMM_TC3X,M,N,K,w.KernelParams(tune=(t0,0,0,0)))
    ctx.launch(k,dA,dB,dC); got=dC.read(np.float32).reshape(M,N)
    ref=A.astype(np.float64)@B.astype(np.float64)
    print(M,N,K,t0,k.name, float(np.abs(got-ref).max()), flush=True)
" 2>&1 | tail -13
done
export B200MM_TC3X_SPLIT=2
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for e in 0 1 2; do echo "SPLIT=$e"; export B200MM_TC3X_SPLIT=$e; timeout 200 python tools/size_sweep.py 2>&1 | head -12; done
