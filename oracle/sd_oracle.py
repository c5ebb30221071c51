"""TEST INFRASTRUCTURE ONLY -- CPU oracle of the SD-backend beam-search step (BASELINE.json config 5).

Plain-PyTorch functional restatement (no nn.Module, weights from a state dict) of
  * `UNet2DConditionModel.forward` for the SD-1.5 family: sd/diffusers/src/diffusers/models/unets/
    unet_2d_condition.py:1039-1310 (time embedding :1098-1116, down :1192-1224, mid :1241-1260, up :1276-1302,
    post-process :1305-1308), blocks unet_2d_blocks.py (CrossAttnDownBlock2D, DownBlock2D, UNetMidBlock2DCrossAttn,
    UpBlock2D, CrossAttnUpBlock2D), resnet.py ResnetBlock2D.forward, downsampling.py / upsampling.py,
    transformers/transformer_2d.py Transformer2DModel.forward, attention.py BasicTransformerBlock.forward,
    attention_processor.py AttnProcessor2_0, activations.py GEGLU, embeddings.py get_timestep_embedding;
  * the reference's edited `DDIMScheduler.step` (eta defaults to 1, returns (prev_sample, pred_original_sample)):
    sd/diffusers/src/diffusers/schedulers/scheduling_ddim.py:342-471, `_get_variance` :218-227,
    `set_timesteps` :297-340 ('leading' spacing + steps_offset);
  * the beam loop of `StableDiffusionPipeline.__call__`: pipelines/stable_diffusion/pipeline_stable_diffusion.py:1045-1170,
    with an injectable decode stage (identity for config 5: the 4-channel latent takes the non-RGB branch of
    sd/scorers.py:66-67 after the uint8 quantisation of :1115).

Pinned by oracle/make_golden_sd.py against the vendored diffusers modules themselves (tests/golden/sd_*.pt).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module."""
import math

import numpy as np
from typing import Callable, Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F


# ---------------------------------------------------------------------------------------------- UNet
def sd_config_from_state_dict(sd: Dict[str, torch.Tensor]) -> dict:
    """Recover (block_out_channels, layers_per_block, which blocks carry attention, cross dim) from key names/shapes."""
    n_down = 1 + max(int(k.split('.')[1]) for k in sd if k.startswith('down_blocks.'))
    boc = [sd[f'down_blocks.{i}.resnets.0.conv1.weight'].shape[0] for i in range(n_down)]
    lpb = 1 + max(int(k.split('.')[3]) for k in sd if k.startswith('down_blocks.0.resnets.'))
    cross = [f'down_blocks.{i}.attentions.0.norm.weight' in sd for i in range(n_down)]
    cdim = next(v.shape[1] for k, v in sd.items() if k.endswith('attn2.to_k.weight'))
    return dict(block_out_channels=boc, layers_per_block=lpb, cross_attn_down=cross, cross_attention_dim=cdim, heads=8,
                groups=32, in_channels=sd['conv_in.weight'].shape[1], out_channels=sd['conv_out.weight'].shape[0])


def timestep_embedding(t: torch.Tensor, dim: int) -> torch.Tensor:
    """embeddings.py get_timestep_embedding with flip_sin_to_cos=True, downscale_freq_shift=0 (SD-1.5 config)."""
    half = dim // 2
    exponent = -math.log(10000) * torch.arange(0, half, dtype=torch.float32, device=t.device)
    exponent = exponent / (half - 0)
    emb = t[:, None].float() * torch.exp(exponent)[None, :]
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    return torch.cat([emb[:, half:], emb[:, :half]], dim=-1)


def _resnet(sd, p, x, temb_act, groups=32, eps=1e-5):
    """resnet.py ResnetBlock2D.forward (time_embedding_norm='default', output_scale_factor=1)."""
    h = F.silu(F.group_norm(x, groups, sd[f'{p}.norm1.weight'], sd[f'{p}.norm1.bias'], eps))
    h = F.conv2d(h, sd[f'{p}.conv1.weight'], sd[f'{p}.conv1.bias'], padding=1)
    t = F.linear(temb_act, sd[f'{p}.time_emb_proj.weight'], sd[f'{p}.time_emb_proj.bias'])[:, :, None, None]
    h = h + t
    h = F.silu(F.group_norm(h, groups, sd[f'{p}.norm2.weight'], sd[f'{p}.norm2.bias'], eps))
    h = F.conv2d(h, sd[f'{p}.conv2.weight'], sd[f'{p}.conv2.bias'], padding=1)
    if f'{p}.conv_shortcut.weight' in sd:
        x = F.conv2d(x, sd[f'{p}.conv_shortcut.weight'], sd[f'{p}.conv_shortcut.bias'])
    return x + h


def _attention(sd, p, x, ctx, heads):
    """attention_processor.py AttnProcessor2_0: q/k/v linears without bias, SDPA with scale head_dim^-0.5, to_out.0."""
    B, L, C = x.shape
    kv = x if ctx is None else ctx
    q = F.linear(x, sd[f'{p}.to_q.weight'])
    k = F.linear(kv, sd[f'{p}.to_k.weight'])
    v = F.linear(kv, sd[f'{p}.to_v.weight'])
    hd = C // heads
    q, k, v = [t.view(B, -1, heads, hd).transpose(1, 2) for t in (q, k, v)]
    o = F.scaled_dot_product_attention(q, k, v, attn_mask=None, dropout_p=0.0, is_causal=False)
    o = o.transpose(1, 2).reshape(B, L, C)
    return F.linear(o, sd[f'{p}.to_out.0.weight'], sd[f'{p}.to_out.0.bias'])


def _transformer(sd, p, x, ctx, heads, groups=32):
    """transformer_2d.py Transformer2DModel.forward (continuous input, conv projections, one BasicTransformerBlock)."""
    B, C, H, W = x.shape
    res = x
    h = F.group_norm(x, groups, sd[f'{p}.norm.weight'], sd[f'{p}.norm.bias'], 1e-6)
    h = F.conv2d(h, sd[f'{p}.proj_in.weight'], sd[f'{p}.proj_in.bias'])
    h = h.permute(0, 2, 3, 1).reshape(B, H * W, C)
    t = f'{p}.transformer_blocks.0'
    n = F.layer_norm(h, (C,), sd[f'{t}.norm1.weight'], sd[f'{t}.norm1.bias'], 1e-5)
    h = _attention(sd, f'{t}.attn1', n, None, heads) + h
    n = F.layer_norm(h, (C,), sd[f'{t}.norm2.weight'], sd[f'{t}.norm2.bias'], 1e-5)
    h = _attention(sd, f'{t}.attn2', n, ctx, heads) + h
    n = F.layer_norm(h, (C,), sd[f'{t}.norm3.weight'], sd[f'{t}.norm3.bias'], 1e-5)
    g = F.linear(n, sd[f'{t}.ff.net.0.proj.weight'], sd[f'{t}.ff.net.0.proj.bias'])
    a, gate = g.chunk(2, dim=-1)
    h = F.linear(a * F.gelu(gate), sd[f'{t}.ff.net.2.weight'], sd[f'{t}.ff.net.2.bias']) + h
    h = h.reshape(B, H, W, C).permute(0, 3, 1, 2).contiguous()
    h = F.conv2d(h, sd[f'{p}.proj_out.weight'], sd[f'{p}.proj_out.bias'])
    return h + res


def sd_unet_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, t, ctx: torch.Tensor, cfg: Optional[dict] = None,
                    taps: Optional[dict] = None) -> torch.Tensor:
    """eps = UNet(x [B,4,H,W], timestep t (int or [B]), encoder_hidden_states ctx [B,77,D]).  `taps` (optional dict)
    receives the output of every resnet / transformer / sampler by module name (per-layer parity tests)."""
    cfg = cfg or sd_config_from_state_dict(sd)
    boc, lpb, heads, groups = cfg['block_out_channels'], cfg['layers_per_block'], cfg['heads'], cfg['groups']
    B = x.shape[0]
    tt = torch.as_tensor(t, device=x.device)
    tt = tt.reshape(-1).expand(B) if tt.numel() == 1 else tt
    emb = timestep_embedding(tt, boc[0]).to(x.dtype)
    emb = F.linear(emb, sd['time_embedding.linear_1.weight'], sd['time_embedding.linear_1.bias'])
    emb = F.linear(F.silu(emb), sd['time_embedding.linear_2.weight'], sd['time_embedding.linear_2.bias'])
    temb_act = F.silu(emb)                                               # ResnetBlock2D applies nonlinearity(temb) first

    def tap(name, v):
        if taps is not None:
            taps[name] = v
        return v

    h = tap('conv_in', F.conv2d(x, sd['conv_in.weight'], sd['conv_in.bias'], padding=1))
    skips = [h]
    for i in range(len(boc)):
        for j in range(lpb):
            h = tap(f'down_blocks.{i}.resnets.{j}', _resnet(sd, f'down_blocks.{i}.resnets.{j}', h, temb_act, groups))
            if cfg['cross_attn_down'][i]:
                h = tap(f'down_blocks.{i}.attentions.{j}', _transformer(sd, f'down_blocks.{i}.attentions.{j}', h, ctx, heads, groups))
            skips.append(h)
        if i != len(boc) - 1:
            p = f'down_blocks.{i}.downsamplers.0.conv'
            h = tap(p, F.conv2d(h, sd[f'{p}.weight'], sd[f'{p}.bias'], stride=2, padding=1))
            skips.append(h)
    h = tap('mid_block.resnets.0', _resnet(sd, 'mid_block.resnets.0', h, temb_act, groups))
    h = tap('mid_block.attentions.0', _transformer(sd, 'mid_block.attentions.0', h, ctx, heads, groups))
    h = tap('mid_block.resnets.1', _resnet(sd, 'mid_block.resnets.1', h, temb_act, groups))
    cross_up = list(cfg['cross_attn_down'])[::-1]
    for i in range(len(boc)):
        for j in range(lpb + 1):
            h = torch.cat([h, skips.pop()], dim=1)
            h = tap(f'up_blocks.{i}.resnets.{j}', _resnet(sd, f'up_blocks.{i}.resnets.{j}', h, temb_act, groups))
            if cross_up[i]:
                h = tap(f'up_blocks.{i}.attentions.{j}', _transformer(sd, f'up_blocks.{i}.attentions.{j}', h, ctx, heads, groups))
        if i != len(boc) - 1:
            p = f'up_blocks.{i}.upsamplers.0.conv'
            h = F.interpolate(h, scale_factor=2.0, mode='nearest')
            h = tap(p, F.conv2d(h, sd[f'{p}.weight'], sd[f'{p}.bias'], padding=1))
    h = F.silu(F.group_norm(h, groups, sd['conv_norm_out.weight'], sd['conv_norm_out.bias'], 1e-5))
    return F.conv2d(h, sd['conv_out.weight'], sd['conv_out.bias'], padding=1)


# ---------------------------------------------------------------------------------------------- DDIM
class DDIMTable:
    """DDIMScheduler(beta_start=.00085, beta_end=.012, 'scaled_linear', clip_sample=False, set_alpha_to_one=False,
    steps_offset=1) after set_timesteps(num_inference_steps): scheduling_ddim.py:190-216, 297-340."""

    def __init__(self, num_inference_steps: int, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012,
                 steps_offset=1):
        betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.final_alpha_cumprod = self.alphas_cumprod[0]                     # set_alpha_to_one=False
        self.num_train_timesteps, self.num_inference_steps = num_train_timesteps, num_inference_steps
        ratio = num_train_timesteps // num_inference_steps                    # 'leading'
        import numpy as np
        ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64) + steps_offset
        self.timesteps = [int(v) for v in ts]

    def coeffs(self, t: int, eta: float = 1.0):
        """(alpha_prod_t, alpha_prod_t_prev, std_dev_t) as fp32 tensors, scheduling_ddim.py:398-435."""
        prev = t - self.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev] if prev >= 0 else self.final_alpha_cumprod
        variance = ((1 - a_prev) / (1 - a_t)) * (1 - a_t / a_prev)                # _get_variance :218-227
        return a_t, a_prev, eta * variance ** 0.5


def ddim_step(tab: DDIMTable, model_output, t: int, sample, variance_noise=None, eta: float = 1.0):
    """scheduling_ddim.py:342-471 (epsilon prediction, no clipping): returns (prev_sample, pred_original_sample)."""
    a_t, a_prev, std = tab.coeffs(t, eta)
    beta_t = 1 - a_t
    pred_x0 = (sample - beta_t ** 0.5 * model_output) / a_t ** 0.5
    direction = (1 - a_prev - std ** 2) ** 0.5 * model_output
    prev = a_prev ** 0.5 * pred_x0 + direction
    if eta > 0 and variance_noise is not None:
        prev = prev + std * variance_noise
    return prev, pred_x0


def latent_brightness(pred_x0: torch.Tensor) -> torch.Tensor:
    """pipeline...:1115 quantisation, then the non-RGB branch of BrightnessScorer (sd/scorers.py:66-67)."""
    u8 = (pred_x0 * 127.5 + 128).clip(0, 255).to(torch.uint8)
    return (u8.float() / 255.0).mean(dim=(1, 2, 3))


def beam_search(unet: Callable, tab: DDIMTable, latents: torch.Tensor, ctx_pair: torch.Tensor, B: int, N: int,
                noises: Sequence[torch.Tensor], guidance: float = 7.5, score_fn: Callable = latent_brightness,
                record: Optional[dict] = None) -> torch.Tensor:
    """pipeline_stable_diffusion.py:1045-1170 with identity decode.  unet(x [M,4,H,W], t, ctx [M,77,D]) -> eps;
    ctx_pair = [uncond, cond] embeddings [2,77,D]; noises[i] = [B, N, 4, H, W] variance noise of step i.
    Returns the best final latent [1,4,H,W]."""
    def guided(x, t):
        xin = torch.cat([x, x])                                                    # :1058 (scale_model_input is identity)
        ctx = torch.cat([ctx_pair[0:1].expand(x.shape[0], -1, -1), ctx_pair[1:2].expand(x.shape[0], -1, -1)])
        e = unet(xin, t, ctx)
        eu, et = e.chunk(2)
        return eu + guidance * (et - eu)                                           # :1073-1075

    beams = [latents.clone() for _ in range(B)]                                    # :1046
    for i, t in enumerate(tab.timesteps):
        cands, scores = [], []
        for bi, beam in enumerate(beams):
            eps = guided(beam, t)
            for n in range(N):
                cand, _ = ddim_step(tab, eps, t, beam, variance_noise=noises[i][bi, n:n + 1])      # :1083
                eps2 = guided(cand, t)                                             # second UNet call at the SAME t (:1090)
                _, pred_next = ddim_step(tab, eps2, t, cand)                       # :1109 (only pred_original_sample is used)
                cands.append(cand)
                scores.append(float(score_fn(pred_next)))
        order = sorted(range(len(scores)), key=lambda k: scores[k], reverse=True)  # stable: lowest flat index wins ties (:1132)
        if record is not None:
            record.setdefault('scores', []).append(torch.tensor(scores))
            record.setdefault('best', []).append(torch.tensor(order[:B]))
        beams = [cands[k] for k in order[:B]]
        if record is not None:
            record.setdefault('beams', []).append(torch.cat(beams))
    best, best_score = beams[0], float('-inf')                                     # :1153-1166
    for cand in beams:
        s = float(score_fn(cand))
        if s > best_score:
            best, best_score = cand, s
    if record is not None:
        record['final_score'] = best_score
    return best


def eps_greedy_search(unet: Callable, tab: DDIMTable, latents: torch.Tensor, ctx_pair: torch.Tensor, N: int, K: int,
                      lam: float, eps: float, method: str, noise: dict, guidance: float = 7.5,
                      score_fn: Callable = latent_brightness, record: Optional[dict] = None):
    """pipeline_stable_diffusion.py:1330-1436 (`eps_greedy` / `zero_order` / `naive`) with identity decode.
    noise['pivot'][i] [1,4,H,W]; noise['dirs'][i][k] [N,4,H,W] (the randn_like draw of candidate n, :1375/:1377);
    noise['r'][i][k][n], noise['u'][i][k][n] = the two torch.rand(1).item() draws (:1373, :1379; u unused on the fresh
    branch).  Returns (final latents, max_score of the last selection or None)."""
    def guided(x, t):
        xin = torch.cat([x, x])                                                    # :1341
        ctx = torch.cat([ctx_pair[0:1].expand(x.shape[0], -1, -1), ctx_pair[1:2].expand(x.shape[0], -1, -1)])
        eu, et = unet(xin, t, ctx).chunk(2)
        return eu + guidance * (et - eu)                                           # :1357-1359

    x = latents.clone()
    max_score = None
    for i, t in enumerate(tab.timesteps):
        eps_pred = guided(x, t)
        pivot = noise['pivot'][i]                                                  # :1366
        if method in ('eps_greedy', 'zero_order'):
            for k in range(K):
                cands = []
                for n in range(N):
                    r = float(noise['r'][i][k][n])
                    if (r < eps) if method == 'eps_greedy' else 0.0:               # :1374 (zero_order never draws fresh)
                        cands.append(noise['dirs'][i][k][n:n + 1])
                    else:
                        to_add = noise['dirs'][i][k][n:n + 1]
                        to_add = to_add / torch.norm(to_add)
                        cands.append(pivot + to_add * float(noise['u'][i][k][n]) * lam * np.sqrt(
                            x.shape[-1] * x.shape[-2] * x.shape[-3]))              # :1379
                scores = []
                for c in cands:
                    lat_c, _ = ddim_step(tab, eps_pred, t, x, variance_noise=c)    # :1384
                    eps2 = guided(lat_c, t)                                        # :1392-1406, the SAME t
                    _, pred_next = ddim_step(tab, eps2, t, lat_c)                  # :1412
                    scores.append(float(score_fn(pred_next)))                      # :1414-1431
                max_score = max(scores)
                best = scores.index(max_score)                                     # first maximal key of the dict (:1435)
                pivot = cands[best]
                if record is not None:
                    record.setdefault('scores', []).append(torch.tensor(scores))
                    record.setdefault('best', []).append(best)
                    record.setdefault('cands', []).append(torch.cat(cands))
        x, _ = ddim_step(tab, eps_pred, t, x, variance_noise=pivot)                # :1437
        if record is not None:
            record.setdefault('x', []).append(x.clone())
    if max_score is None:                                                          # :1469-1474
        max_score = float(score_fn(x))
    return x, max_score


# ---------------------------------------------------------------------------------------------- VAE decode
def vae_decode(sd: Dict[str, torch.Tensor], z: torch.Tensor, taps: Optional[dict] = None, groups: int = 32) -> torch.Tensor:
    """AutoencoderKL._decode (autoencoder_kl.py:287-298) = post_quant_conv + Decoder.forward (vae.py:291-340):
    conv_in -> UNetMidBlock2D (resnet, one-head attention, resnet) -> UpDecoderBlock2D x n (resnets, nearest x2 + conv)
    -> GroupNorm(eps 1e-6) -> SiLU -> conv_out.  z [B,4,h,w] (already divided by scaling_factor) -> image [B,3,8h,8w]."""
    def tap(name, t):
        if taps is not None:
            taps[name] = t
        return t

    def resnet(p, x):                                          # resnet.py ResnetBlock2D.forward, temb = None
        h = F.silu(F.group_norm(x, groups, sd[f'{p}.norm1.weight'], sd[f'{p}.norm1.bias'], 1e-6))
        h = F.conv2d(h, sd[f'{p}.conv1.weight'], sd[f'{p}.conv1.bias'], padding=1)
        h = F.silu(F.group_norm(h, groups, sd[f'{p}.norm2.weight'], sd[f'{p}.norm2.bias'], 1e-6))
        h = F.conv2d(h, sd[f'{p}.conv2.weight'], sd[f'{p}.conv2.bias'], padding=1)
        if f'{p}.conv_shortcut.weight' in sd:
            x = F.conv2d(x, sd[f'{p}.conv_shortcut.weight'], sd[f'{p}.conv_shortcut.bias'])
        return tap(p, x + h)                                   # output_scale_factor = 1

    def attention(p, x):                                       # attention_processor.py AttnProcessor2_0, heads = 1
        B, C, H, W = x.shape
        t = x.view(B, C, H * W)
        t = F.group_norm(t, groups, sd[f'{p}.group_norm.weight'], sd[f'{p}.group_norm.bias'], 1e-6).transpose(1, 2)
        q = F.linear(t, sd[f'{p}.to_q.weight'], sd[f'{p}.to_q.bias'])
        k = F.linear(t, sd[f'{p}.to_k.weight'], sd[f'{p}.to_k.bias'])
        v = F.linear(t, sd[f'{p}.to_v.weight'], sd[f'{p}.to_v.bias'])
        o = F.scaled_dot_product_attention(q[:, None], k[:, None], v[:, None])[:, 0]
        o = F.linear(o, sd[f'{p}.to_out.0.weight'], sd[f'{p}.to_out.0.bias'])
        return tap(p, o.transpose(1, 2).reshape(B, C, H, W) + x)      # residual_connection, rescale_output_factor = 1

    z = F.conv2d(z, sd['post_quant_conv.weight'], sd['post_quant_conv.bias'])
    h = tap('decoder.conv_in', F.conv2d(z, sd['decoder.conv_in.weight'], sd['decoder.conv_in.bias'], padding=1))
    h = resnet('decoder.mid_block.resnets.0', h)
    h = attention('decoder.mid_block.attentions.0', h)
    h = resnet('decoder.mid_block.resnets.1', h)
    i = 0
    while f'decoder.up_blocks.{i}.resnets.0.norm1.weight' in sd:
        j = 0
        while f'decoder.up_blocks.{i}.resnets.{j}.norm1.weight' in sd:
            h = resnet(f'decoder.up_blocks.{i}.resnets.{j}', h)
            j += 1
        p = f'decoder.up_blocks.{i}.upsamplers.0.conv'
        if f'{p}.weight' in sd:                                # upsampling.py Upsample2D: nearest x2, then conv
            h = F.interpolate(h, scale_factor=2.0, mode='nearest')
            h = tap(p, F.conv2d(h, sd[f'{p}.weight'], sd[f'{p}.bias'], padding=1))
        i += 1
    h = F.silu(F.group_norm(h, groups, sd['decoder.conv_norm_out.weight'], sd['decoder.conv_norm_out.bias'], 1e-6))
    return F.conv2d(h, sd['decoder.conv_out.weight'], sd['decoder.conv_out.bias'], padding=1)


def image_brightness(image: torch.Tensor) -> torch.Tensor:
    """pipeline...:1115 quantisation of the decoded image, then BrightnessScorer's RGB branch (sd/scorers.py:43-64)."""
    u8 = (image * 127.5 + 128).clip(0, 255).to(torch.uint8)
    w = torch.tensor([0.2126, 0.7152, 0.0722]).view(1, 3, 1, 1)
    return ((u8.float() / 255.0) * w).sum(dim=1).mean(dim=(1, 2)).clamp(0, 1)


# ---------------------------------------------------------------------------------------------- reference RNG call order
# The search branches draw from torch's GLOBAL generators, in this order (pinned against the real pipeline by
# oracle/make_golden_sd_search.py -> tests/golden/sd_search_tiny.pt, tests/test_sd_oracle.py):
#   * every candidate noise is its own `torch.randn_like(latents)` call of shape [1,4,H,W] (:1080, :1375, :1377);
#   * the scoring-only second `scheduler.step(noise_pred_tminusone, t, latents_cand, **extra_step_kwargs)` (:1109, :1412)
#     runs with eta = 1 and NO variance_noise, so scheduling_ddim.py:457-461 draws -- and discards the effect of -- one
#     more `randn_tensor(model_output.shape, generator=None)` per scored candidate;
#   * eps_greedy / zero_order: `torch.rand(1).item()` decides fresh-vs-perturb (:1373) and, on the perturb branch, scales the
#     direction (:1379) -- on the CPU generator whatever the device, i.e. on a CPU run they interleave with the randn draws.
# A device run (CUDA) keeps the rand(1) draws on the CPU generator and the randn draws on the device generator.
def draw_beam_noise(shape, n_steps: int, B: int, N: int, device=None):
    """The variance noises of the beam branch in the reference's call order: per step, per beam: N candidate draws, then
    the N discarded draws of the scoring-only steps.  Returns noises[i] = [B, N, 4, H, W]."""
    out = []
    for _ in range(n_steps):
        per = []
        for _b in range(B):
            cands = [torch.randn(shape, device=device) for _ in range(N)]                  # :1080
            for _ in range(N):
                torch.randn(shape, device=device)                                          # :1109 -> scheduling_ddim.py:457
            per.append(torch.cat(cands))
        out.append(torch.stack(per))
    return out


def draw_eps_greedy_noise(shape, n_steps: int, N: int, K: int, eps: float, method: str, device=None) -> dict:
    """pivot / r / dirs / u of the eps_greedy, zero_order and naive branches in the reference's call order (:1366-1379)
    plus the discarded per-candidate draw of the scoring-only step (:1412)."""
    noise = dict(pivot=[], dirs=[], r=[], u=[])
    for _ in range(n_steps):
        noise['pivot'].append(torch.randn(shape, device=device))                            # :1366
        dk, rk, uk = [], [], []
        if method in ('eps_greedy', 'zero_order'):
            for _k in range(K):
                dirs, rs, us = [], [], []
                for _n in range(N):
                    r = torch.rand(1).item()                                                # :1373 (CPU generator)
                    rs.append(r)
                    dirs.append(torch.randn(shape, device=device))                          # :1375 / :1377
                    fresh = (r < eps) if method == 'eps_greedy' else 0.0
                    us.append(0.0 if fresh else torch.rand(1).item())                       # :1379
                for _n in range(N):
                    torch.randn(shape, device=device)                                       # :1412 -> scheduling_ddim.py:457
                dk.append(torch.cat(dirs))
                rk.append(rs)
                uk.append(us)
        noise['dirs'].append(dk)
        noise['r'].append(rk)
        noise['u'].append(uk)
    return noise


def mcts_as_shipped(unet: Callable, tab: DDIMTable, latents: torch.Tensor, ctx_pair: torch.Tensor, N: int, S: int,
                    guidance: float = 7.5, score_fn: Callable = latent_brightness, device=None):
    """The `mcts` branch AS SHIPPED (pipeline_stable_diffusion.py:1172-1333).  Nothing in it ever scores a rollout or
    increments `visits` / `total_reward`, hence: the selection walk never leaves the root (:1211: a child with visits == 0
    stops it), the first N of the S iterations expand one child each with a fresh `randn_like` (:1243), every iteration runs
    a rollout to the last timestep whose result is discarded, and `max(children, key = -inf for all)` (:1306) returns the
    FIRST child.  The observable behaviour is therefore: x <- DDIM step with the step's first noise draw -- plus the RNG
    consumption of the discarded work (one child draw per expansion, one `randn_tensor` per rollout step: the rollout calls
    scheduler.step with eta = 1 and no variance_noise, :1288 -> scheduling_ddim.py:457).  This restatement skips the UNet
    evaluations whose outputs are unused and reproduces the draws; oracle/make_golden_sd_search.py pins it (final latents,
    max_score) against the real pipeline.  Returns (latents, max_score)."""
    def guided(x, t):
        xin = torch.cat([x, x])
        ctx = torch.cat([ctx_pair[0:1].expand(x.shape[0], -1, -1), ctx_pair[1:2].expand(x.shape[0], -1, -1)])
        eu, et = unet(xin, t, ctx).chunk(2)
        return eu + guidance * (et - eu)

    x = latents.clone()
    T = len(tab.timesteps)
    for i, t in enumerate(tab.timesteps):
        first = None
        for s in range(S):
            if s < N:                                                              # expansion (:1216-1251)
                noise = torch.randn(x.shape, device=device)
                first = noise if first is None else first
            for _ in range(i, T):                                                  # rollout (:1276-1300): discarded draws
                torch.randn(x.shape, device=device)
        if first is not None:                                                      # :1305-1308
            x, _ = ddim_step(tab, guided(x, t), t, x, variance_noise=first)
    return x, float(score_fn(x))                                                   # :1466-1471
