import sys, ctypes as C, numpy as np
sys.path.insert(0,'.')
import torch
from phantomsdr_b200 import SpectrumConfig, AM, USB, LSB
from phantomsdr_b200.backend import B200FFT
from phantomsdr_b200.synth import make_clients
cfg=SpectrumConfig(sps=35_000_000, fft_size=1<<20)
n=cfg.audio_fft_size; F=8
eng=B200FFT(cfg.fft_size,1,cfg.downsample_levels,0,0); eng.set_output_additional_size(n); eng.plan_c2c()
eng.set_hop_ring(16); eng.set_batch_frames(F); eng.clients_create(1024,n,12000)
for i,c in enumerate(make_clients(cfg,1024,modes=(AM,USB,LSB))): eng.client_open(i,c.l,c.mid,c.r,c.mode)
ring=torch.as_tensor(eng.device_hop_ring(16),device='cuda'); ring.normal_(0,1e-3)
fn=0
for it in range(6):
    eng.execute_device(0,F); eng.clients_execute_device(fn,F); fn+=F
eng.sync()
out=(C.c_longlong*8)()
eng.L.b200_debug_tail_profile(eng.h,1,None)
reps=10
for it in range(reps):
    eng.execute_device(0,F); eng.clients_execute_device(fn,F); fn+=F
eng.sync()
eng.L.b200_debug_tail_profile(eng.h,0,out)
v=np.array(list(out)[:7],float)/(reps*F)
print("cycles per frame per phase [load,sum1,avg,sum2,peak,gain,store]:",v.round(0), "total",v.sum())
