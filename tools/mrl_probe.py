import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lightretriever_b200 as lr
dev="cuda"
N,d,Q,k=1_100_000,4096,10000,100
c=torch.nn.functional.normalize(torch.randn(N,d,device=dev),dim=-1).bfloat16()
q=torch.nn.functional.normalize(torch.randn(Q,d,device=dev),dim=-1).bfloat16()
def t(fn,it=5):
    for _ in range(3): fn()
    torch.cuda.synchronize(); a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b)/it
for m in (128, 512, 1024):
    cs=(1.0/c[:,:m].float().norm(dim=1)).contiguous(); qs=(1.0/q[:,:m].float().norm(dim=1)).contiguous()
    cc=c[:,:m].contiguous(); qq=q[:,:m].contiguous()
    for name,env in [("default",{}),("narrow",{"LR_FLATIP_WIDE":"0"}),("noepi",{"LR_FLATIP_DEBUG":"1"}),("wide_multicast",{"LR_FLATIP_CLUSTER":"2"}),("wide_norefresh",{"LR_FLATIP_REFRESH":"0"})]:
        for k_ in ("LR_FLATIP_DEBUG","LR_FLATIP_PREFIX_DOCS","LR_FLATIP_CLUSTER","LR_FLATIP_SCHED","LR_FLATIP_TEAM_WINDOW","LR_FLATIP_REFRESH","LR_FLATIP_REFRESH_GROWTH","LR_FLATIP_WIDE"): os.environ.pop(k_,None)
        os.environ.update(env)
        a=t(lambda: lr.flatip_topk(q,c,k,d_used=m,q_scale=qs,c_scale=cs))
        b=t(lambda: lr.flatip_topk(qq,cc,k))
        print(json.dumps({"m":m,"cfg":name,"ms_fullwidth_scaled":round(a,2),"ms_compact_unscaled":round(b,2)}),flush=True)
