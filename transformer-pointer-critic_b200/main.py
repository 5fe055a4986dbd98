"""Train-then-test entry point with the reference's signature and on-disk artefacts (main.py:40-83):

    results/<env name>/<start date>/config.json        every config used (main.py:62)
    results/<env name>/<start date>/env.txt             the profile table (env.py:443-444)
    results/<env name>/<start date>/<folder>/logs.csv   training curves (agents/plotter.py:19,99-126)
    results/<env name>/<start date>/model/actor.npz     actor weights (agents/trainer.py:178-184)
    results/<env name>/<start date>/tests/test.csv      tester statistics (misc/csv_writer.py:3-35)

The matplotlib learning / reward plots (agents/plotter.py:21-97) are not produced: the CSV holds the same series.

    python -m tpc_b200.main [env_name] [n_iterations]
"""
from __future__ import annotations

import json
import os
import sys
from datetime import datetime

from .agent import Agent
from .configs import get_configs
from .env import env_factory
from .trainer import trainer

LOG_DIR = "./results/"


def log_training_stats(data, location, file_name):
    """agents/plotter.py:99-126 (same header, column order and %.3f formatting)."""
    (average_rewards_buffer, min_rewards_buffer, max_rewards_buffer, value_loss_buffer, bins_policy_loss_buffer,
     bins_total_loss_buffer, bins_entropy_buffer) = data
    os.makedirs(location, exist_ok=True)
    with open(f"{location}/{file_name}.csv", "w") as fp:
        fp.write("Step;Avg Reward;Max Reward;Min Reward;Value Loss;Bin Entropy;Total Bin Loss;Bin Policy Loss\n")
        for index in range(len(average_rewards_buffer)):
            fp.write(f"{index};{average_rewards_buffer[index]:.3f};{max_rewards_buffer[index]:.3f};"
                     f"{min_rewards_buffer[index]:.3f};{value_loss_buffer[index]:.3f};"
                     f"{bins_entropy_buffer[index]:.3f};{bins_total_loss_buffer[index]:.3f};"
                     f"{bins_policy_loss_buffer[index]:.3f}\n")


def training_plotter(data, env, agent, agent_config: dict, trainer_config: dict, log_dir: str):
    """agents/plotter.py:5-19: only the CSV export."""
    if not trainer_config["export_stats"]["export_stats"]:
        return
    location = os.path.join(log_dir, trainer_config["export_stats"]["folder"])
    log_training_stats(data, location, "logs")


def runner(env_type="custom", env_name="ResourceV3", agent_name="tpc", device="cuda:0", log_root=LOG_DIR,
           overrides: dict | None = None):
    """main.py:40-83. `overrides` = {"trainer_config": {...}, "env_config": {...}, "tester_config": {...}} is applied on
    top of configs/<env_name>.json (the reference edits the JSON file instead)."""
    start_date = datetime.now().replace(microsecond=0).isoformat()
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        # one log directory per job: rank 0's timestamp (every rank stamping its own would scatter config.json / env.txt)
        stamp = [start_date]
        dist.broadcast_object_list(stamp, src=0)
        start_date = stamp[0]
    agent_config, trainer_config, env_config, tester_config, _, all_configs = get_configs(env_name, agent_name)
    for section, target in (("trainer_config", trainer_config), ("env_config", env_config),
                            ("tester_config", tester_config)):
        target.update((overrides or {}).get(section, {}))
    env, tester = env_factory(env_type, env_name, env_config, device)
    agent_config = env.add_stats_to_agent_config(agent_config)
    agent = Agent("transformer", agent_config, device, num_features=env.num_features)
    log_dir = os.path.join(log_root, env.name, start_date)
    os.makedirs(log_dir, exist_ok=True)
    json.dump(all_configs, open(os.path.join(log_dir, "config.json"), "w"), indent=6)
    env.store_dataset(os.path.join(log_dir, "env.txt"))
    print("Training...")
    training_history = trainer(env, agent, trainer_config, True, log_dir)
    print("\nTraining Done...")
    training_plotter(training_history, env, agent, agent_config, trainer_config, log_dir)
    print("\nTesting...")
    result = tester(env, agent, tester_config, log_dir)
    print("\nEnd... Goodbye!")
    return log_dir, training_history, result


if __name__ == "__main__":  # pragma: no cover
    name = sys.argv[1] if len(sys.argv) > 1 else "ResourceV3"
    ov = {"trainer_config": {"n_iterations": int(sys.argv[2])}} if len(sys.argv) > 2 else None
    runner(env_name=name, overrides=ov)
