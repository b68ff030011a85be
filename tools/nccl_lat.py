import os, torch, torch.distributed as dist, time
rank=int(os.environ["RANK"]); local=int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
dist.init_process_group("nccl"); dev=torch.device("cuda",local)
buf=torch.zeros(1380,device=dev); big=torch.empty(256<<20,dtype=torch.uint8,device=dev); work=torch.randn(4096,4096,device=dev)
def t(fn,n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize(); dist.barrier(device_ids=[local]); torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n
a=t(lambda: dist.all_reduce(buf))
b=t(lambda: (work@work, dist.all_reduce(buf)))
c=t(lambda: (work@work,))
d=t(lambda: (big.fill_(1), work@work, dist.all_reduce(buf)))
e=t(lambda: (big.fill_(1), work@work))
if rank==0: print(f"allreduce alone {a*1e3:.1f} us | matmul+allreduce {b*1e3:.1f} us | matmul {c*1e3:.1f} us | fill+mm+ar {d*1e3:.1f} | fill+mm {e*1e3:.1f}")
dist.destroy_process_group()
