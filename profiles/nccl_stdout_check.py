"""Does anything but our own line reach stdout when a NCCL process group starts? (bench.py prints ONE JSON line.)"""
import os, sys, torch, torch.distributed as dist
os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')
if os.environ.get('NCCL_DEBUG', '').upper() in ('VERSION', 'WARN'):
    del os.environ['NCCL_DEBUG']
os.environ.setdefault('MASTER_ADDR','127.0.0.1'); os.environ.setdefault('MASTER_PORT','29533')
print('NCCL_DEBUG env =', os.environ.get('NCCL_DEBUG'), file=sys.stderr)
dist.init_process_group('nccl', rank=0, world_size=1, device_id=torch.device('cuda:0'))
t=torch.ones(4,device='cuda'); dist.all_reduce(t); torch.cuda.synchronize()
print('{"ok": 1}')
dist.destroy_process_group()
