"""Policy-in-the-loop evaluation on the device: the deterministic actor of a Stable-Baselines3
``MultiInputPolicy`` (the models the reference trains and ships: main.py:36-55,
scripts/common/results/models_backup/<env>/<env>_<ALGO>/model.zip) driving ``BlueSkyVectorEnv`` without the
observations ever leaving the GPU.

Only the actor's forward pass is re-stated (SB3 itself is not needed):
  * SAC          ``actor.latent_pi.{0,2}`` (ReLU) -> ``actor.mu`` -> tanh
  * TD3 / DDPG   ``actor.mu.{0,2,4}`` (ReLU, ReLU, tanh)
  * PPO / A2C    ``mlp_extractor.policy_net.{0,2}`` (tanh) -> ``action_net`` -> clip to [-1, 1]
SB3's ``CombinedExtractor`` concatenates the flattened Dict observation in the key order of the gymnasium
``Dict`` space, which sorts the keys of a plain ``dict`` (the reference passes plain dicts, e.g.
horizontal_cr_env.py:49-60): the actor input is the observation in ALPHABETICAL key order, not in declaration
order.  ``SB3Actor`` applies that permutation to the env's flat device observation.

The MLP runs through torch (cuBLAS): it is the caller's policy, not part of the simulator path.
"""
import io
import json
import zipfile
from collections import OrderedDict

import numpy as np
import torch

_LAYOUTS = (
    ("actor.latent_pi.0.weight", (("actor.latent_pi.0", "relu"), ("actor.latent_pi.2", "relu"), ("actor.mu", "tanh"))),
    ("actor.mu.0.weight", (("actor.mu.0", "relu"), ("actor.mu.2", "relu"), ("actor.mu.4", "tanh"))),
    ("mlp_extractor.policy_net.0.weight", (("mlp_extractor.policy_net.0", "tanh"), ("mlp_extractor.policy_net.2", "tanh"),
                                           ("action_net", "clip"))),
)


def extract_actor(state_dict):
    """SB3 ``policy.pth`` state dict -> [(W, b, activation), ...] of the deterministic actor."""
    for probe, layers in _LAYOUTS:
        if probe in state_dict:
            return [(np.asarray(state_dict[n + ".weight"], dtype=np.float32), np.asarray(state_dict[n + ".bias"], dtype=np.float32), a)
                    for n, a in layers]
    raise ValueError("unrecognised SB3 policy layout: " + ", ".join(list(state_dict)[:6]))


def read_sb3_zip(path):
    """(actor layers, metadata) from an SB3 ``model.zip``."""
    z = zipfile.ZipFile(path)
    sd = torch.load(io.BytesIO(z.read("policy.pth")), map_location="cpu", weights_only=True)
    meta = json.loads(z.read("data"))
    return extract_actor({k: v.numpy() for k, v in sd.items()}), meta


def save_actor_npz(path, layers, obs_keys):
    out = {"activations": np.array([a for _, _, a in layers]), "obs_keys": np.array(list(obs_keys))}
    for i, (w, b, _) in enumerate(layers):
        out[f"W{i}"], out[f"b{i}"] = w, b
    np.savez_compressed(path, **out)


def load_actor_npz(path):
    g = np.load(path)
    acts = [str(a) for a in g["activations"]]
    return [(g[f"W{i}"], g[f"b{i}"], a) for i, a in enumerate(acts)], [str(k) for k in g["obs_keys"]]


class SB3Actor(torch.nn.Module):
    """Deterministic SB3 actor evaluated on the env's flat device observation ``[E, obs_dim]``."""

    def __init__(self, layers, obs_layout, device):
        super().__init__()
        self.acts = [a for _, _, a in layers]
        self.weights = torch.nn.ParameterList([torch.nn.Parameter(torch.as_tensor(w), requires_grad=False) for w, _, _ in layers])
        self.biases = torch.nn.ParameterList([torch.nn.Parameter(torch.as_tensor(b), requires_grad=False) for _, b, _ in layers])
        cols = []
        for k in sorted(obs_layout):                     # gymnasium Dict: plain-dict keys are sorted
            off, w = obs_layout[k][0], obs_layout[k][1]
            cols.extend(range(off, off + w))
        assert len(cols) == layers[0][0].shape[1], (len(cols), layers[0][0].shape)
        self.register_buffer("perm", torch.as_tensor(cols, dtype=torch.long))
        self.to(device)

    @classmethod
    def from_zip(cls, path, venv):
        layers, _ = read_sb3_zip(path)
        return cls(layers, venv.obs_layout, venv.device)

    @classmethod
    def from_npz(cls, path, venv):
        layers, _ = load_actor_npz(path)
        return cls(layers, venv.obs_layout, venv.device)

    @torch.no_grad()
    def forward(self, flat_obs):
        x = flat_obs.index_select(1, self.perm)
        for w, b, a in zip(self.weights, self.biases, self.acts):
            x = torch.addmm(b, x, w.t())
            x = torch.relu(x) if a == "relu" else (torch.tanh(x) if a == "tanh" else torch.clamp(x, -1.0, 1.0))
        return x


@torch.no_grad()
def evaluate(venv, actor, episodes_per_env=1, max_steps=None):
    """Runs ``actor`` (None = uniform random actions) until every env finished ``episodes_per_env`` episodes.
    Everything stays on the device; returns a dict of CPU arrays: per-episode ``returns`` and ``lengths`` plus the
    mean of every info key at episode end (the columns of the reference's CSV logs, utils/logger.py:18-33)."""
    assert venv.autoreset_mode == "same_step", "evaluate() relies on same-step autoreset"
    E, dev = venv.num_envs, venv.device
    venv.reset_torch()
    obs = venv.t["obs"]
    ret = torch.zeros(E, device=dev)
    length = torch.zeros(E, device=dev)
    left = torch.full((E,), episodes_per_env, device=dev, dtype=torch.int32)
    rets, lens, infos = [], [], []
    cap = max_steps or (venv.cfg.max_episode_steps * episodes_per_env + 1 if venv.cfg.max_episode_steps else 100000)
    g = torch.Generator(device=dev).manual_seed(0)
    for _ in range(cap):
        a = actor(obs) if actor is not None else torch.rand((E, venv.layout.act_dim), device=dev, generator=g) * 2 - 1
        _, rew, term, trunc = venv.step_torch(a)
        obs = venv.t["obs"]
        active = left > 0
        ret += torch.where(active, rew, torch.zeros_like(rew))
        length += active.float()
        done = ((term != 0) | (trunc != 0)) & active
        if bool(done.any()):
            rets.append(ret[done].cpu())
            lens.append(length[done].cpu())
            infos.append(venv.t["info"][:, done].T.cpu())
            ret = torch.where(done, torch.zeros_like(ret), ret)
            length = torch.where(done, torch.zeros_like(length), length)
            left = left - done.int()
        if not bool((left > 0).any()):
            break
    out = OrderedDict(returns=torch.cat(rets).numpy() if rets else np.zeros(0),
                      lengths=torch.cat(lens).numpy() if lens else np.zeros(0))
    if infos:
        inf = torch.cat(infos).numpy()
        for i, k in enumerate(venv.spec_b200.info_keys):
            out["info_" + k] = inf[:, i]
    return out
