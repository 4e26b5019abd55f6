"""PPO samples/s of the GPU-resident trainer (BASELINE.json configs 0, 3, 4 shapes).

    python profiles/measure_ppo.py                       # 1 GPU: configs 0 (num_envs=4), vs_dummy wide, selfplay
    torchrun --nproc-per-node 8 profiles/measure_ppo.py --multi   # config 4: sharded rollout + NCCL grad all-reduce

Prints one JSON line per run (rank 0). samples/s = env transitions consumed by PPO per second of
wall clock, rollout and updates included (the reference's README quotes 100k steps in 10-20 min
at num_envs=4, i.e. 83-167 samples/s)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from inversus_b200.sharding import dist_env  # noqa: E402
from inversus_b200.training import train  # noqa: E402

KEYS = ("n_gpus", "num_envs", "steps", "steps_per_env", "batch_size", "precision", "elapsed_s", "samples_per_s",
        "rollout_s", "update_s", "rollout_env_steps_per_s", "steady_state", "episodes", "win_rate", "policy_loss", "value_loss", "entropy")


def run(name, **kw):
    torch.manual_seed(0)
    out = train(log_dir=f"/tmp/inv_measure_{name}", seed=0, quiet=True, save=False, **kw)
    if dist_env()[0] == 0:
        print(json.dumps({"run": name, **{k: out.get(k) for k in KEYS}}), flush=True)


if __name__ == "__main__":
    if "--config4" in sys.argv:  # BASELINE.json configs[4] literally: 8 x 1 048 576 envs
        world = dist_env()[2]
        n = (1 << 20) * world
        run(f"config4_vs_dummy_hard_{world}gpu_{n}envs", mode="vs_dummy", num_envs=n, total_steps=n * 8 * 2,
            opponent_difficulty="hard", rollout_steps=8, batch_size=32768, epochs=1, precision="bf16")
        if torch.distributed.is_initialized():
            torch.distributed.destroy_process_group()
    elif "--multi" in sys.argv:
        world = dist_env()[2]
        n = 131072 * world
        run(f"vs_dummy_hard_{world}gpu_sharded", mode="vs_dummy", num_envs=n, total_steps=n * 16 * 4,
            opponent_difficulty="hard", rollout_steps=16, batch_size=32768, epochs=1, precision="bf16")
        if torch.distributed.is_initialized():
            torch.distributed.destroy_process_group()
    else:
        run("config0_vs_dummy_envs4_fp32_refgae", mode="vs_dummy", num_envs=4, total_steps=8192,
            opponent_difficulty="hard", precision="fp32", reference_gae=True)
        run("config0_vs_dummy_envs4_bf16", mode="vs_dummy", num_envs=4, total_steps=8192,
            opponent_difficulty="hard", precision="bf16")
        run("vs_dummy_hard_envs65536_bf16", mode="vs_dummy", num_envs=65536, total_steps=65536 * 16 * 3,
            opponent_difficulty="hard", rollout_steps=16, batch_size=16384, precision="bf16")
        run("config3_selfplay_envs65536_bf16", mode="selfplay", num_envs=65536, total_steps=65536 * 16 * 3,
            rollout_steps=16, batch_size=16384, precision="bf16")
