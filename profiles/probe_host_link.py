import sys, json, torch
sys.path.insert(0,'.')
from rigidbody_simulation_b200 import shard
dev=torch.device("cuda:0")
pinned=torch.empty(58720256//8, dtype=torch.float64).pin_memory()
print(json.dumps(shard.host_link_bandwidth(dev, pinned)))
