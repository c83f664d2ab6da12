import sys, ctypes as C, numpy as np
sys.path.insert(0,'.')
import torch
from phantomsdr_b200 import SpectrumConfig, AM, USB, LSB
from phantomsdr_b200.backend import B200FFT
from phantomsdr_b200.synth import make_clients
cfg=SpectrumConfig(sps=35_000_000, fft_size=1<<20)
n=cfg.audio_fft_size; F=int(sys.argv[1]) if len(sys.argv)>1 else 64
eng=B200FFT(cfg.fft_size,1,cfg.downsample_levels,0,0); eng.set_output_additional_size(n); eng.plan_c2c()
eng.set_hop_ring(max(16,F+1)); eng.set_batch_frames(F)
BANKS=int(sys.argv[2]) if len(sys.argv)>2 else 1  # 2: the forward group of batch k+1 runs beside the clients of batch k, as in a bench step
if BANKS>1: eng.set_pipeline(BANKS)
eng.clients_create(1024,n,12000)
for i,c in enumerate(make_clients(cfg,1024,modes=(AM,USB,LSB))): eng.client_open(i,c.l,c.mid,c.r,c.mode)
ring=torch.as_tensor(eng.device_hop_ring(max(16,F+1)),device='cuda'); ring.normal_(0,1e-3)
fn=0
bn=0
def one():
    global fn,bn
    if BANKS>1: eng.select_bank(bn%BANKS)
    eng.execute_device(0,F); eng.clients_execute_device(fn,F); fn+=F; bn+=1
for it in range(6): one()
eng.sync()
out=(C.c_longlong*64)()
eng.L.b200_debug_tail_profile(eng.h,1,None)
reps=10
for it in range(reps): one()
eng.sync()
eng.L.b200_debug_tail_profile(eng.h,0,out)
v=np.array(list(out)[:45],float).reshape(15,3)/(reps*F)
names=["load","sum1","sum2","block","gain","peak0","peak1","peak2","peak3","out0","out1","out2","out3","suffix0","suffix1"]
print("cycles per frame (CTA 0): stage, waiting for input, waiting for ring space, total")
for nme,row in zip(names,v): print(f"  {nme:7s} {row[0]:9.0f} {row[1]:9.0f} {row[2]:9.0f}   busy {row[2]-row[0]-row[1]:9.0f}")
