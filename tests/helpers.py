"""Shared test helpers: golden loading, seeded nets/noise (mirrors oracle/make_golden.py)."""
import os

import torch

from oracle import edm_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name), weights_only=False)


def oracle_net(cfg, seed):
    spec = O.build_unet_spec(**cfg)
    sd = O.seeded_state_dict(O.unet_param_shapes(spec), seed)
    return O.OracleNet(spec, sd), spec, sd


def search_inputs(gold):
    """Regenerate latents/labels/noise exactly as oracle/make_golden.py:gen_search drew them."""
    cfg, seed = gold['cfg'], gold['seed']
    b, N, K, num_steps, method = gold['b'], gold['N'], gold['K'], gold['num_steps'], gold['method']
    g = torch.Generator().manual_seed(seed + 2)
    res, c = cfg['img_resolution'], cfg['in_channels']
    latents = torch.randn(b, c, res, res, generator=g)
    labels = torch.eye(cfg['label_dim'])[torch.randint(cfg['label_dim'], (b,), generator=g)] if cfg['label_dim'] else None
    pre = {}
    if method == 'EPS_GREEDY':
        for i in range(num_steps):
            pre[f'pivot_{i}'] = torch.randn(b, c, res, res, generator=g, dtype=torch.float64)
            pre[i] = torch.randn(b, K, N, c, res, res, generator=g, dtype=torch.float64)
            if gold['eps'] == 1.0:
                for k in range(K):
                    for n in range(N):
                        pre[f'fresh_{i}_{k}_{n}'] = torch.randn(b, c, res, res, generator=g, dtype=torch.float64)
    elif method == 'REJECTION_SAMPLING':
        for i in range(num_steps):
            pre[i] = torch.randn(b, N, c, res, res, generator=g, dtype=torch.float64)
    return latents, labels, pre


def scale_fn_from(gold):
    table = gold['scales']
    return lambda i, k, n: table[f'{i}_{k}_{n}']
