"""ImageNet classifier scorer on a B200 vs the reference's own outputs
(tests/golden/classifier_*.pt from oracle/make_golden.py).  Tolerance: bf16 torso, fp32 head; the score is
a softmax probability, compared relatively (measured ~1e-2 relative on random-init weights)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import classifier_oracle as CO  # noqa: E402
from tests.helpers import load_golden  # noqa: E402


@pytest.fixture(scope='module')
def cls():
    from diffusion_tts_b200 import build
    build.build()
    import diffusion_tts_b200.classifier as c
    return c


@pytest.mark.parametrize('name', ['classifier_tiny.pt', 'classifier_full.pt'])
def test_classifier_matches_reference_golden(cls, name):
    g = load_golden(name)
    sd = CO.seeded_classifier_state_dict(CO.classifier_param_shapes(**g['cfg']), g['seed'])
    scorer = cls.ImageNetScorer(sd, device='cuda')
    M = g['images'].shape[0]
    scores = scorer(g['images'].cuda(), g['labels'].cuda(), torch.zeros(M, device='cuda')).cpu()
    fp = scorer.engine.plan(M)
    logits = fp.logits.cpu()
    rel = ((logits - g['logits']).norm() / g['logits'].norm()).item()
    print(name, 'logits rel err', rel, 'scores', scores.tolist(), 'ref', g['scores'].tolist())
    assert rel < 3e-2
    assert torch.allclose(scores, g['scores'], rtol=0.1, atol=1e-5)
    assert scores[0] == scores[1] or not torch.equal(g['labels'][0], g['labels'][1])     # identical images
    assert torch.equal(fp.logits[0], fp.logits[1])                                       # -> identical bits


def test_imagenet_scorer_in_search_loop(cls):
    """BASELINE.json configs[3] in miniature: eps_greedy with the classifier scorer in the hot loop
    (generic scorer protocol: uint8 Tweedie images are materialised on the GPU and scored there)."""
    import diffusion_tts_b200.denoiser as den
    import diffusion_tts_b200.edm.main as em
    from oracle import edm_oracle as O
    from tests.helpers import oracle_net, search_inputs
    g = load_golden('search_eps_greedy_tiny.pt')
    gc = load_golden('classifier_tiny.pt')
    csd = CO.seeded_classifier_state_dict(CO.classifier_param_shapes(**gc['cfg']), gc['seed'])
    onet, spec, sd = oracle_net(g['cfg'], g['seed'])
    latents, labels, pre = search_inputs(g)
    net = den.B200Denoiser(sd, device='cuda')
    table = den.StepTable(net, 'cuda', g['num_steps'], **g['sampler_kw'])
    scorer = cls.ImageNetScorer(csd, device='cuda')
    params = em.SamplingParams(N=g['N'], K=1, eps=0.0, lambda_param=0.15, scorer=scorer)
    seen = []

    class Spy:
        def __call__(self, im, lab, t):
            s = scorer(im, lab, t)
            seen.append((im.cpu(), lab.cpu(), s.cpu()))
            return s
    params.scorer = Spy()
    x, rec = em.eps_greedy_search(net, latents.cuda(), labels.cuda(), params, table,
                                  precomputed_noise={k: v.cuda() for k, v in pre.items()}, record=True, escalate=False)
    assert len(seen) == g['num_steps']
    for im, lab, s in seen:                      # every scored batch agrees with the CPU oracle's classifier
        ref = CO.imagenet_score(csd, gc['cfg'], im, lab, torch.zeros(im.shape[0]))
        assert torch.allclose(s, ref, rtol=0.1, atol=1e-5)
