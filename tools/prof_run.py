"""Two forwards of the ECAPA B=256 backbone (for ncu captures: first = warm-up, second = profiled)."""
import sys
sys.path.insert(0, '.')
import torch
import __graft_entry__ as ge
ge.build()
from mvector.models import build_model
from mvector.utils.utils import dict_to_object
from oracle import models as om
from loguru import logger
logger.remove()
margs = dict(embd_dim=192, pooling_type='ASP', channels=[512, 512, 512, 512, 1536])
m = build_model(80, dict_to_object({'model_conf': {'model': 'EcapaTdnn', 'model_args': margs}}))
m.load_state_dict(om.random_state_dict('EcapaTdnn', 80, seed=0, **margs))
B, T = 256, 298
prog = m.program(B, T)
f = torch.randn(B, T, 80, device='cuda') * 2
e = torch.empty(B, 192, device='cuda')
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
for _ in range(n):
    prog.run(f, e)
torch.cuda.synchronize()
print('done', float(e.abs().sum()))
