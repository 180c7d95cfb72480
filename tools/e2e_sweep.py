"""End-to-end host pipeline (pinned H2D -> fused kernel -> D2H) against chunk size, stream count and graph replay: python tools/e2e_sweep.py"""
import sys, time, torch
sys.path.insert(0,'/root/repo')
from skiing_analysis_pytorch_b200 import api, synth
dev=torch.device('cuda:0'); T,J=1_000_000,17
clip=synth.make_clip('2b',T,J,seed=0)
h_k=torch.from_numpy(clip.x_vm).pin_memory()
h_X=torch.empty((T,J,3),dtype=torch.float32).pin_memory(); h_s=torch.empty((T,2,4),dtype=torch.float32).pin_memory()
for graph, ramp in ((False, 0), (False, 4096), (False, 8192), (False, 16384), (True, 8192)):
  for chunk in (32768,65536,131072,262144):
    for ns in (2,3,4):
        kw=dict(graph=graph,ramp_from=ramp,K=clip.K,R=clip.R,t=clip.t,dist=synth.DIST_CALIB,want=("X","stats"),chunk_frames=chunk,n_streams=ns)
        for _ in range(2): api.triangulate_reproject_host(h_k,out={"X":h_X,"stats":h_s},**kw)
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(8): api.triangulate_reproject_host(h_k,out={"X":h_X,"stats":h_s},**kw)
        e1.record(); torch.cuda.synchronize()
        print(f"graph {graph} ramp {ramp} chunk {chunk:7d} streams {ns}: {e0.elapsed_time(e1)/8:.3f} ms/step", flush=True)
