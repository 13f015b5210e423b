import numpy as np
from .tenalg import multi_mode_dot


def random_tucker(shape, rank, full=False, orthogonal=False, random_state=None, **kw):
    rng = np.random.RandomState(random_state)
    if isinstance(rank, int):
        rank = [rank] * len(shape)
    factors = [rng.random_sample((s, r)) for s, r in zip(shape, rank)]
    core = rng.random_sample(tuple(rank))
    if full:
        return multi_mode_dot(core, factors)
    return core, factors
