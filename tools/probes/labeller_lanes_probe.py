import os, sys, time, torch
sys.path.insert(0, '/root/repo')
from mulactseg_b200 import labeller, synth
dev='cuda:0'
for name,(h,w,nseg,c,rho,bs) in {"voc":(375,500,150,21,0.3,16),"city":(1024,2048,2048,20,0.08,4)}.items():
    feats=synth.features(bs,256,h,w,seed=10,device=dev); logits=synth.logits(bs,c,h,w,"normal",seed=30,device=dev,coherent=4)
    spx=synth.superpixel_map(1,h,w,nseg,"jitter",seed=3,device=dev).repeat(bs,1,1); trg=synth.multihot_targets(1,nseg,c,seed=4,device=dev,p_ignore=0.0).repeat(bs,1,1)
    mask=synth.region_mask(spx,nseg,rho,seed=5)
    f=lambda: labeller.pseudo_label_generation(None,feats,logits,trg,mask,spx,check=False)
    for _ in range(3): f()
    torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(10): f()
    t_host=(time.perf_counter()-t0)/10/bs*1e3; torch.cuda.synchronize(); t_all=(time.perf_counter()-t0)/10/bs*1e3
    print(name, os.environ.get("MAS_LABELLER_LANES"), f"host enqueue {t_host:.4f} ms/img, total {t_all:.4f} ms/img", flush=True)
