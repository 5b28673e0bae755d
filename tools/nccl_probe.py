import os, torch, torch.distributed as dist, datetime, time, sys
r=int(os.environ['RANK']); torch.cuda.set_device(r)
t=time.time()
dist.init_process_group('nccl', device_id=torch.device('cuda',r), timeout=datetime.timedelta(seconds=60))
x=torch.full((4,),float(r),device='cuda'); out=torch.empty(8,device='cuda')
dist.all_gather_into_tensor(out,x); torch.cuda.synchronize()
print('rank',r,'allgather ok',out.tolist(),'%.1fs'%(time.time()-t)); sys.stdout.flush()
dist.barrier(); dist.destroy_process_group()
