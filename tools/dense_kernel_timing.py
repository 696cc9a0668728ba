"""CUDA-event timings of the dense kernels of one GASFM block at d = 256 (weight gradient single / batched, forward projections,
concatenated input gradient in its 3xTF32 / fp16 two-pass / fp16 one-pass forms).  Usage: python tools/dense_kernel_timing.py [E]
(A/B switches: GASFM_WGRAD_LEAN=0, GASFM_GEMM_DEBUG=32.)"""
import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from gasfm_b200 import ops
dev='cuda:0'
torch.manual_seed(0)
E=int(sys.argv[1]) if len(sys.argv)>1 else 495592
x=torch.relu(torch.randn(E,256,device=dev)); dys=[torch.randn(E,256,device=dev) for _ in range(3)]
amax3=torch.stack([d.abs().max() for d in dys]); ax=x.abs().max().reshape(1)
r={}
r['wgrad_single_ms']=bench.timed_batches(lambda: ops.wgrad_f16x2(dys[0],x,amax3[:1],ax))
r['wgrad_multi3_ms']=bench.timed_batches(lambda: ops.wgrad_f16x2_multi(dys,x,amax3,ax))
dw,db=ops.wgrad_f16x2_multi(dys,x,amax3,ax)
ref=dys[1].double().t()@x.double()
r['err']=float((dw[1].double()-ref).abs().max()/ref.abs().max())
print(json.dumps(r))
w=torch.randn(256,256,device=dev)/16
r2={}
r2['gemm_single_ms']=bench.timed_batches(lambda: ops.gemm_f16x2(x,w))
r2['gemm_3groups_ms']=bench.timed_batches(lambda: ops.gemm_f16x2_groups(x,[w,w,w],[None,None,None]))
wc=torch.cat([w,w,w],1)
r2['dx_tf32_ms']=bench.timed_batches(lambda: ops.gemm_tf32x3_cat(dys,wc))
r2['dx_f16_ms']=bench.timed_batches(lambda: ops.gemm_f16x2_cat(dys,wc))
print(json.dumps(r2))
rms=[d.abs().amax(1) for d in dys]
r3={'dx_f16_rowmax_ms': bench.timed_batches(lambda: ops.gemm_f16x2_cat(dys,wc,rowmax=rms))}
print(json.dumps(r3))
