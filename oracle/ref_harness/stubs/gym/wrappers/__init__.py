from gym.wrappers.time_limit import TimeLimit
