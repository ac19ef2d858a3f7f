"""Logging surface of the reference (utils/logging_utils.py:42-136, wandb keys of run.py:303-325), host-side only.

``CSVLogger`` writes the two files the reference writes -- ``<policy_path>_results.csv`` (Hours, Step, then
SuccessRate / EpisodeLength / Return per env) and ``<policy_path>_losses.csv`` (Hours, Step, TD Error, Grad Norm and the six
Q / target statistics) -- with the same headers, the same row layout and the same "do not overwrite on resume" rule, so
plotting scripts written against the reference read them unchanged.  ``get_logger`` returns it for ``--disable-wandb`` and
the ``wandb`` module (initialised with the reference's project / group naming) otherwise; without wandb installed it falls
back to the CSV files with a one-line notice.
"""
import csv
import os
from datetime import datetime
from typing import Dict

LOSS_COLUMNS = (("TD Error", "losses/TD_Error"), ("Grad Norm", "losses/Grad_Norm"), ("Max Q Value", "losses/Max_Q_Value"),
                ("Mean Q Value", "losses/Mean_Q_Value"), ("Min Q Value", "losses/Min_Q_Value"),
                ("Max Target Value", "losses/Max_Target_Value"), ("Mean Target Value", "losses/Mean_Target_Value"),
                ("Min Target Value", "losses/Min_Target_Value"))
GROUP_KEYS = ["model", "obs_embed", "a_embed", "in_embed", "context", "layers", "bag_size", "gate", "identity", "history",
              "pos"]                                                           # logging_utils.py:116-128


def timestamp() -> str:
    return datetime.now().strftime("%B %d, %H:%M:%S")                           # logging_utils.py:27-28


def loss_log_values(agent, hours: float) -> Dict[str, float]:
    """The ``losses/*`` entries of one log call (run.py:303-313) from the agent's RunningAverage statistics."""
    return {"losses/TD_Error": agent.td_errors.mean(), "losses/Grad_Norm": agent.grad_norms.mean(),
            "losses/Max_Q_Value": agent.qvalue_max.mean(), "losses/Mean_Q_Value": agent.qvalue_mean.mean(),
            "losses/Min_Q_Value": agent.qvalue_min.mean(), "losses/Max_Target_Value": agent.target_max.mean(),
            "losses/Mean_Target_Value": agent.target_mean.mean(), "losses/Min_Target_Value": agent.target_min.mean(),
            "losses/hours": hours}


class CSVLogger:
    """``log(results, step)`` has the signature of ``wandb.log`` (logging_utils.py:42-107)."""

    def __init__(self, path: str, args):
        self.results_path = path + "_results.csv"
        self.losses_path = path + "_losses.csv"
        self.envs = list(args.envs)
        if not os.path.exists(self.results_path):              # a resumed run appends to the files it already has
            head = ["Hours", "Step"]
            for env in self.envs:
                head += [f"{env}/SuccessRate", f"{env}/EpisodeLength", f"{env}/Return"]
            with open(self.results_path, "w") as f:
                csv.writer(f).writerow(head)
        if not os.path.exists(self.losses_path):
            with open(self.losses_path, "w") as f:
                csv.writer(f).writerow(["Hours", "Step"] + [c for c, _ in LOSS_COLUMNS])

    def log(self, results: Dict[str, float], step: int) -> None:
        row = [results["losses/hours"], step]
        for env in self.envs:
            row += [results[f"{env}/SuccessRate"], results[f"{env}/EpisodeLength"], results[f"{env}/Return"]]
        with open(self.results_path, "a") as f:
            csv.writer(f).writerow(row)
        with open(self.losses_path, "a") as f:
            csv.writer(f).writerow([results["losses/hours"], step] + [results[k] for _, k in LOSS_COLUMNS])


def get_logger(policy_path: str, args, wandb_kwargs: Dict[str, str]):
    """logging_utils.py:110-136."""
    if getattr(args, "disable_wandb", True):
        return CSVLogger(policy_path, args)
    try:
        import wandb
    except ImportError:
        print("[dtqn_b200] wandb is not installed; logging to CSV files instead", flush=True)
        return CSVLogger(policy_path, args)
    config = vars(args)
    wandb.init(project=config["project_name"],
               group="_".join(f"{k}={v}" for k, v in config.items() if k in GROUP_KEYS), config=config, **wandb_kwargs)
    return wandb
