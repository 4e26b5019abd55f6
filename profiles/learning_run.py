"""Learning-quality evidence: train PPO on the GPU-resident loop and log the reference's own
metrics (training_log.csv: step, episode, avg_reward, win_rate, avg_ep_len, losses). The reference
ships two such logs (runs/*/training_log.csv: 0.35 win rate vs the easy dummy after 200 k steps,
0.57 vs the hard dummy after 500 k more); this produces ours for the same opponents.

    python profiles/learning_run.py easy|hard|selfplay [total_updates]
"""
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from inversus_b200.training import train  # noqa: E402

difficulty = sys.argv[1] if len(sys.argv) > 1 else "easy"
updates = int(sys.argv[2]) if len(sys.argv) > 2 else 40
num_envs, rollout_steps = 16384, 64
torch.manual_seed(0)
log_dir = f"/tmp/inv_learning_{difficulty}"
mode = "selfplay" if difficulty == "selfplay" else "vs_dummy"
out = train(mode, num_envs=num_envs, total_steps=num_envs * rollout_steps * updates, log_dir=log_dir,
            opponent_difficulty="easy" if mode == "selfplay" else difficulty, precision="bf16", rollout_steps=rollout_steps, batch_size=8192,
            epochs=4, lr=1e-4, seed=0, quiet=False, save=False)
dst = os.path.join(ROOT, "gpurun_out", f"learning_{difficulty}_training_log.csv")
os.makedirs(os.path.dirname(dst), exist_ok=True)
shutil.copy(os.path.join(log_dir, "training_log.csv"), dst)
print(json.dumps({k: out[k] for k in ("steps", "episodes", "elapsed_s", "samples_per_s", "win_rate", "avg_reward",
                                      "avg_ep_len", "wins_per_kstep")}))
