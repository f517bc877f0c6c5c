"""Developer tool (torchrun under gpurun --gpus N): BASELINE config 2 -- EcoDispatch, 8 192 envs per GPU."""
import os, sys; sys.path.insert(0, '.')
import torch, torch.distributed as dist
world, rank, lr = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    sys.stdout.flush(); saved = os.dup(1); os.dup2(2, 1)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr)); dist.barrier(); torch.cuda.synchronize()
    os.dup2(saved, 1); os.close(saved)
from opfgym_b200 import envs
B = 8192
env = envs.EcoDispatch(num_envs=B, train_data="full_uniform", test_data="full_uniform", n_profile_steps=672,
                       rank=rank, world_size=world, device=torch.device("cuda", lr), seed=1, copy_outputs=False)
env.reset(seed=1)
a = torch.rand(B, env.single_action_space.shape[0], dtype=torch.float64, device="cuda")
for _ in range(5): env.step(a)
env.reset_statistics()
if world > 1: dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
K = 30
for _ in range(K): env.step(a)
e1.record(); torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
if world > 1: dist.all_reduce(ms, op=dist.ReduceOp.MAX)
st = env.episode_statistics(reduce=True)
if rank == 0:
    print(f"config2 EcoDispatch {world} GPU x {B} envs: {ms.item()/K:.3f} ms/step -> {world*B*K/(ms.item()*1e-3):.4e} env-steps/s "
          f"conv={st['converged_share']:.4f} iters={st['mean_iterations']:.2f}")
