#!/usr/bin/env python
"""PPO on the batched engine with observations, actions and rollout buffers resident on the GPU (BASELINE config 5:
16384 single-car envs on martinsville, Discrete(5)).  The shape of the reference's learn/ppo.py loop -- collect n_steps
from a vectorised env, then clipped-objective updates -- with `SubprocVecEnv([...] * 8)` replaced by one
`NascarVectorEnv` whose `step_torch` takes and returns CUDA tensors, so nothing crosses PCIe inside the loop.

    python examples/ppo_rollout.py --envs 16384 --track martinsville --discrete 1 --iters 5

Multi-GPU: launch with torchrun; each rank owns its own envs (no data-path collective).  --sync gather (default): once per
iteration the rollouts and episode statistics of every rank are gathered on the learner rank (rank 0) over NCCL
(nascargymnasium_b200.distributed.gather_rollout / reduce_stats), rank 0 runs the clipped-objective updates on the whole
batch and broadcasts the new weights.  --sync allreduce: every rank updates on its own rollouts and the gradients are
averaged with one all-reduce per minibatch.  Prints one JSON line per iteration from rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nascargymnasium_b200.vector_env import NascarVectorEnv  # noqa: E402


class ActorCritic(nn.Module):
    """MlpPolicy-sized network (2 x 64 tanh) with a categorical or diagonal-Gaussian head."""

    def __init__(self, obs_dim: int, n_actions: int, discrete: bool):
        super().__init__()
        self.discrete = discrete
        self.pi = nn.Sequential(nn.Linear(obs_dim, 64), nn.Tanh(), nn.Linear(64, 64), nn.Tanh(), nn.Linear(64, n_actions))
        self.v = nn.Sequential(nn.Linear(obs_dim, 64), nn.Tanh(), nn.Linear(64, 64), nn.Tanh(), nn.Linear(64, 1))
        self.log_std = nn.Parameter(torch.zeros(n_actions))

    def dist(self, obs):
        out = self.pi(obs)
        if self.discrete:
            return torch.distributions.Categorical(logits=out, validate_args=False)
        return torch.distributions.Normal(out, self.log_std.exp(), validate_args=False)

    def act(self, obs):
        d = self.dist(obs)
        # (Normal.sample() validates its scale with a device->host sync, which a CUDA graph capture cannot contain)
        a = d.sample() if self.discrete else d.loc + d.scale * torch.randn_like(d.loc)
        logp = d.log_prob(a) if self.discrete else d.log_prob(a).sum(-1)
        return a, logp, self.v(obs).squeeze(-1)

    def evaluate(self, obs, a):
        d = self.dist(obs)
        logp = d.log_prob(a) if self.discrete else d.log_prob(a).sum(-1)
        ent = d.entropy() if self.discrete else d.entropy().sum(-1)
        return logp, ent, self.v(obs).squeeze(-1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=16384)
    ap.add_argument("--track", default="martinsville")
    ap.add_argument("--discrete", type=int, default=1)
    ap.add_argument("--n-steps", type=int, default=128)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--epochs", type=int, default=2)
    ap.add_argument("--minibatches", type=int, default=4)
    ap.add_argument("--gamma", type=float, default=0.99)
    ap.add_argument("--lam", type=float, default=0.95)
    ap.add_argument("--clip", type=float, default=0.2)
    ap.add_argument("--ent-coef", type=float, default=0.002)        # learn/ppo.py:97
    ap.add_argument("--lr", type=float, default=3e-4)
    ap.add_argument("--sync", default="gather", choices=["gather", "allreduce"], help="multi-GPU: gather rollouts to the learner "
                    "rank and broadcast weights, or average gradients")
    ap.add_argument("--graph", type=int, default=1, help="1: capture one rollout step (policy + env kernel + bookkeeping) in a CUDA "
                    "graph and replay it, instead of launching ~60 small kernels per step from Python")
    args = ap.parse_args()

    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    torch.manual_seed(1 + rank)

    E, T, discrete = args.envs, args.n_steps, bool(args.discrete)
    track_file = None if args.track == "all" else f"tracks/{args.track}.track"
    # (a captured step cannot contain the host-side re-grouping of envs that moved to another track: with --graph the envs of a
    # --track all run keep the track they drew at reset)
    venv = NascarVectorEnv(E, track_file=track_file, discrete_action_space=discrete, device=local, redraw_tracks=not args.graph)
    obs = venv.reset_torch()
    net = ActorCritic(38, 5 if discrete else 2, discrete).to(dev)
    opt = torch.optim.Adam(net.parameters(), lr=args.lr, eps=1e-5)

    b_obs = torch.empty((T, E, 38), device=dev)
    b_act = torch.empty((T, E), dtype=torch.int64, device=dev) if discrete else torch.empty((T, E, 2), device=dev)
    b_logp, b_val, b_rew, b_done = (torch.empty((T, E), device=dev) for _ in range(4))
    ep_ret = torch.zeros(E, device=dev)
    done_returns, done_count = torch.zeros((), device=dev), torch.zeros((), device=dev)

    # static tensors so that one rollout step can be captured once and replayed; `obs` is a view of the engine's output buffer
    t_idx = torch.zeros((), dtype=torch.int64, device=dev)

    def rollout_step(t):
        """policy -> env -> bookkeeping for time index t (a Python int, or the device scalar t_idx inside the graph)."""
        a, logp, val = net.act(obs)
        if isinstance(t, int):
            b_obs[t], b_act[t], b_logp[t], b_val[t] = obs, a, logp, val
        else:
            b_obs.index_copy_(0, t.view(1), obs.unsqueeze(0)); b_act.index_copy_(0, t.view(1), a.unsqueeze(0))
            b_logp.index_copy_(0, t.view(1), logp.unsqueeze(0)); b_val.index_copy_(0, t.view(1), val.unsqueeze(0))
        env_a = a.to(torch.int32) if discrete else a.clamp(-1.0, 1.0)
        _obs, rew, te, tr, final = venv.step_torch(env_a)           # writes the engine's buffers in place: _obs is `obs`
        done = (te | tr).to(torch.float32)
        # an episode cut by the 180 s limit is not a terminal state: bootstrap from the value of its last observation, as
        # SB3's on-policy collector does for "TimeLimit.truncated" (final rows of envs that did not finish are stale and
        # masked out; `rew` is a fresh tensor so the engine's buffer is left alone)
        trunc_only = (tr != 0) & (te == 0)
        rew = rew + torch.where(trunc_only, args.gamma * net.v(final).squeeze(-1), torch.zeros_like(rew))
        if isinstance(t, int):
            b_rew[t], b_done[t] = rew, done
        else:
            b_rew.index_copy_(0, t.view(1), rew.unsqueeze(0)); b_done.index_copy_(0, t.view(1), done.unsqueeze(0))
        ep_ret.add_(rew)
        done_returns.add_((ep_ret * done).sum())
        done_count.add_(done.sum())
        ep_ret.mul_(1.0 - done)

    def build_step_graph():
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s), torch.no_grad():
            for _ in range(3):                                      # warm-up on the side stream (allocator, lazy init)
                rollout_step(t_idx)
        torch.cuda.current_stream().wait_stream(s)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g), torch.no_grad():
            rollout_step(t_idx)
        return g

    graph = None
    for it in range(args.iters):
        # ---------------------------------------------------------------- rollout: policy and env both on the device
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.no_grad():
            if args.graph:
                if graph is None:
                    graph = build_step_graph()
                for t in range(T):
                    t_idx.fill_(t)
                    graph.replay()
            else:
                for t in range(T):
                    rollout_step(t)
            last_val = net.v(obs).squeeze(-1)
        torch.cuda.synchronize()
        t_roll = time.perf_counter() - t0
        # ---------------------------------------------------------------- GAE + clipped-objective updates
        t0 = time.perf_counter()
        with torch.no_grad():
            adv = torch.zeros((T, E), device=dev)
            last = torch.zeros(E, device=dev)
            for t in reversed(range(T)):
                nxt = last_val if t == T - 1 else b_val[t + 1]
                nonterm = 1.0 - b_done[t]
                delta = b_rew[t] + args.gamma * nxt * nonterm - b_val[t]
                last = delta + args.gamma * args.lam * nonterm * last
                adv[t] = last
            ret = adv + b_val
        g_obs, g_act, g_logp, g_adv, g_ret = b_obs, b_act, b_logp, adv, ret
        learner = True
        if world > 1 and args.sync == "gather":
            # the per-iteration exchange: (T, E_local, ...) tensors concatenated along the env axis on the learner rank
            from nascargymnasium_b200 import distributed as D
            g_obs, g_act, g_logp, g_adv, g_ret = (D.gather_rollout(x, dst=0) for x in (b_obs, b_act, b_logp, adv, ret))
            learner = rank == 0
        loss_v = loss_p = torch.zeros((), device=dev)
        if learner:
            fo, fa = g_obs.reshape(-1, 38), g_act.reshape(-1) if discrete else g_act.reshape(-1, 2)
            fl, fadv, fret = g_logp.reshape(-1), g_adv.reshape(-1), g_ret.reshape(-1)
            n = fo.shape[0]
        for _ in range(args.epochs if learner else 0):
            perm = torch.randperm(n, device=dev)
            for mb in perm.chunk(args.minibatches):
                logp, ent, v = net.evaluate(fo[mb], fa[mb])
                a_mb = fadv[mb]
                a_mb = (a_mb - a_mb.mean()) / (a_mb.std() + 1e-8)
                ratio = (logp - fl[mb]).exp()
                loss_p = -torch.min(ratio * a_mb, ratio.clamp(1 - args.clip, 1 + args.clip) * a_mb).mean()
                loss_v = 0.5 * (v - fret[mb]).pow(2).mean()
                loss = loss_p + 0.5 * loss_v - args.ent_coef * ent.mean()
                opt.zero_grad(set_to_none=True)
                loss.backward()
                if world > 1 and args.sync == "allreduce":
                    for p_ in net.parameters():
                        if p_.grad is not None:              # (log_std has no gradient with a categorical head)
                            dist.all_reduce(p_.grad, op=dist.ReduceOp.SUM)
                            p_.grad /= world
                nn.utils.clip_grad_norm_(net.parameters(), 0.5)
                opt.step()
        if world > 1 and args.sync == "gather":
            for p_ in net.parameters():
                dist.broadcast(p_.data, src=0)
            from nascargymnasium_b200 import distributed as D
            tot = D.reduce_stats({"done_count": float(done_count), "done_returns": float(done_returns)}, device=dev)
        else:
            tot = {"done_count": float(done_count), "done_returns": float(done_returns)}
        torch.cuda.synchronize()
        t_upd = time.perf_counter() - t0
        if rank == 0:
            print(json.dumps({"iter": it, "envs_per_gpu": E, "n_gpus": world, "n_steps": T,
                              "rollout_env_steps_per_s": world * E * T / t_roll, "rollout_s": t_roll, "update_s": t_upd,
                              "mean_step_reward": float(b_rew.mean()), "episodes_done": tot["done_count"],
                              "mean_episode_return": tot["done_returns"] / max(tot["done_count"], 1.0), "sync": args.sync if world > 1 else None,
                              "policy_loss": float(loss_p.detach()), "value_loss": float(loss_v.detach()), "obs_device": str(obs.device)}), flush=True)
    venv.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
