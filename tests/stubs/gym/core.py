class Env:
    metadata = {"render.modes": []}
    reward_range = (-float("inf"), float("inf"))
    spec = None
    action_space = None
    observation_space = None

    def step(self, action): raise NotImplementedError
    def reset(self): raise NotImplementedError
    def render(self, mode="human"): raise NotImplementedError
    def close(self): pass
    def seed(self, seed=None): return

    @property
    def unwrapped(self):
        return self

    def __enter__(self): return self
    def __exit__(self, *a): self.close(); return False


class GoalEnv(Env):
    pass


class Wrapper(Env):
    def __init__(self, env):
        self.env = env
        self.action_space = self.env.action_space
        self.observation_space = self.env.observation_space
        self.reward_range = self.env.reward_range
        self.metadata = self.env.metadata

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError("attempted to get missing private attribute '{}'".format(name))
        return getattr(self.env, name)

    @property
    def spec(self): return self.env.spec
    @classmethod
    def class_name(cls): return cls.__name__
    def step(self, action): return self.env.step(action)
    def reset(self, **kwargs): return self.env.reset(**kwargs)
    def render(self, mode="human", **kwargs): return self.env.render(mode, **kwargs)
    def close(self): return self.env.close()
    def seed(self, seed=None): return self.env.seed(seed)
    def compute_reward(self, achieved_goal, desired_goal, info): return self.env.compute_reward(achieved_goal, desired_goal, info)

    @property
    def unwrapped(self): return self.env.unwrapped


class ObservationWrapper(Wrapper):
    def reset(self, **kwargs): return self.observation(self.env.reset(**kwargs))
    def step(self, action):
        observation, reward, done, info = self.env.step(action)
        return self.observation(observation), reward, done, info
    def observation(self, observation): raise NotImplementedError


class RewardWrapper(Wrapper):
    def reset(self, **kwargs): return self.env.reset(**kwargs)
    def step(self, action):
        observation, reward, done, info = self.env.step(action)
        return observation, self.reward(reward), done, info
    def reward(self, reward): raise NotImplementedError


class ActionWrapper(Wrapper):
    def reset(self, **kwargs): return self.env.reset(**kwargs)
    def step(self, action): return self.env.step(self.action(action))
    def action(self, action): raise NotImplementedError
    def reverse_action(self, action): raise NotImplementedError
