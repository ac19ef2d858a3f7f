"""Top-level entry with the reference's ``python run.py ...`` spelling; see dtqn_b200/run.py."""
from dtqn_b200.run import get_args, run_experiment

if __name__ == "__main__":
    run_experiment(get_args())
