"""Base-class shim for the un-vendored third-party dependency ``unioncom==0.4.0``.

The reference's ``JAMIE`` subclasses ``unioncom.UnionCom.UnionCom`` (jamie/jamie.py:17-23, 29, 111) but on the hot path
only uses its constructor (attribute defaults) and ``unioncom.utils.init_random_seed`` (jamie/jamie.py:142).  Both are
restated here from the published package; PARITY UNPINNED for the defaults (no reference test covers them; the values
visible in the reference's stored notebook logs -- epoch_pd=2000, "use random seed: 666" -- agree).
"""
import random

import torch


class UnionCom(object):
    def __init__(self, integration_type='MultiOmics', epoch_pd=2000, epoch_DNN=100, epsilon=0.01, lr=0.001,
                 batch_size=100, rho=10, beta=1, perplexity=30, log_DNN=10, log_pd=100, manual_seed=666, delay=0,
                 kmax=40, output_dim=32, distance_mode='geodesic', project_mode='tsne'):
        self.integration_type = integration_type
        self.epoch_pd = epoch_pd
        self.epoch_DNN = epoch_DNN
        self.epsilon = epsilon
        self.lr = lr
        self.batch_size = batch_size
        self.rho = rho
        self.beta = beta
        self.perplexity = perplexity
        self.log_DNN = log_DNN
        self.log_pd = log_pd
        self.manual_seed = manual_seed
        self.delay = delay
        self.kmax = kmax
        self.output_dim = output_dim
        self.distance_mode = distance_mode
        self.project_mode = project_mode


def init_random_seed(manual_seed):
    """Seeds python ``random`` and torch (NOT numpy), printing the seed like the original."""
    seed = random.randint(1, 10000) if manual_seed is None else manual_seed
    print("use random seed: {}".format(seed))
    random.seed(seed)
    torch.manual_seed(seed)
    return seed
