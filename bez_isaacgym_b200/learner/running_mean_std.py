"""``RunningMeanStd`` with rl_games' interface and state layout (rl_games/algos_torch/running_mean_std.py, v1.1.3;
used by the reference at ``bez_isaacgym/utils/players.py:38,71-72`` and checkpointed as ``running_mean_std`` /
``reward_mean_std``): fp64 buffers ``running_mean``, ``running_var``, ``count`` (so reference checkpoints load),
``forward(input, unnorm=False)``, statistics updated in training mode only.

Kernels: one pass of pivoted fp64 moments (block -> grid, fixed order), a one-block merge with the reference's
parallel-variance update, one streaming normalise pass.  With a process group the pivoted sums of all ranks are
SUM-all-reduced between moments and merge, which makes the update EXACT across shards."""
import torch
from torch import nn

from .. import dist as bdist
from .. import ops


class RunningMeanStd(nn.Module):
    def __init__(self, insize, epsilon=1e-05, per_channel=False, norm_only=False, process_group=None):
        super().__init__()
        if per_channel or norm_only:
            raise NotImplementedError("per_channel / norm_only are not used by the BezKick configuration")
        if isinstance(insize, (tuple, list)):
            if len(insize) != 1:
                raise NotImplementedError("only 1-D observation shapes")
            insize = insize[0]
        self.insize = int(insize)
        self.epsilon = float(epsilon)
        self.process_group = process_group
        self.register_buffer("running_mean", torch.zeros(self.insize, dtype=torch.float64))
        self.register_buffer("running_var", torch.ones(self.insize, dtype=torch.float64))
        self.register_buffer("count", torch.ones((), dtype=torch.float64))
        self._acc = None
        self._scratch = None
        self._pivot = None
        self._acc_ext = None

    def _workspace(self, device):
        if self._acc is None or self._acc.device != device:
            c = self.insize
            self._acc = torch.empty(1 + 2 * c, dtype=torch.float64, device=device)
            self._scratch = torch.empty(ops.rms_scratch_doubles(c), dtype=torch.float64, device=device)
            self._pivot = torch.empty(c, dtype=torch.float64, device=device)
            self._acc_ext = torch.empty(2 + 4 * c, dtype=torch.float64, device=device)

    def update(self, x: torch.Tensor):
        """Merge the batch moments of ``x`` into the running statistics.  ``x``: (m, insize) contiguous, or a slab view
        ``obses[:, e0:e0+E]`` of time-major rollout storage (read in place, see ``learner.experience.SlabDataset``)."""
        self._workspace(x.device)
        self._pivot.copy_(self.running_mean)              # replicated on all ranks -> identical pivots
        if x.is_contiguous():
            ops.rms_moments(x.view(-1, self.insize), self._pivot, self._acc, self._scratch)
        else:
            ops.rms_moments_slabs(x, self._pivot, self._acc, self._scratch)
        bdist.allreduce_sum_(self._acc, self.process_group)
        ops.rms_merge(self._acc, self._pivot, self.running_mean, self.running_var, self.count.view(1))

    # ------------------------------------------------------------------ planned updates (one moment pass per distinct batch)
    def plan(self, batches, order):
        """rl_games updates the obs normaliser with the SAME minibatches in every mini-epoch (``calc_gradients``:
        ``obs = self.running_mean_std(obs)`` in train mode), and the batch moments do not depend on the policy.  ``plan`` takes
        the epoch's distinct ``batches`` (contiguous (m, insize) tensors or slab views) and the update ``order`` (indices into
        ``batches``, e.g. ``[0, 1, 2, 3] * mini_epochs``): ONE moments pass per distinct batch (pivot = the running mean now), one
        SUM all-reduce of all of them when distributed, one kernel that replays the reference's merge for every update in
        order.  Afterwards ``running_*`` hold the state after the last update and ``planned(u, x)`` normalises with the
        statistics after update ``u`` -- exactly what ``forward`` would have used at that point of the epoch."""
        if not self.training:
            raise RuntimeError("plan() replays train-mode updates: call .train() first")
        dev = batches[0].device
        self._workspace(dev)
        c, nb = self.insize, len(batches)
        key = (tuple(int(o) for o in order), nb, str(dev))
        if getattr(self, "_plan_key", None) != key:
            if min(key[0]) < 0 or max(key[0]) >= nb:
                raise ValueError("order must index into batches")
            self._plan_order = torch.tensor(key[0], dtype=torch.int32, device=dev)
            self._plan_acc = torch.empty(nb, 1 + 2 * c, dtype=torch.float64, device=dev)
            self._plan_seq = torch.empty(len(key[0]), 2, c, dtype=torch.float64, device=dev)
            self._plan_key = key
        self._pivot.copy_(self.running_mean)              # replicated on all ranks -> identical pivots
        if self._equally_spaced(batches):                 # minibatches = consecutive env blocks of one rollout tensor: one launch pair
            ops.rms_moments_slabs_batched([x.detach() for x in batches], self._pivot, self._plan_acc, self._scratch)
        else:
            for b, x in enumerate(batches):
                x = x.detach()
                if x.is_contiguous():
                    ops.rms_moments(x.view(-1, c), self._pivot, self._plan_acc[b], self._scratch)
                else:
                    ops.rms_moments_slabs(x, self._pivot, self._plan_acc[b], self._scratch)
        bdist.allreduce_sum_(self._plan_acc, self.process_group)
        ops.rms_merge_sequence(self._plan_acc, self._plan_order, self._pivot, self.running_mean, self.running_var, self.count.view(1),
                               self._plan_seq)
        return self._plan_seq

    @staticmethod
    def _equally_spaced(batches):
        """Equally spaced views of one tensor with one geometry (what the batched kernels take in a single launch)."""
        nb = len(batches)
        step = batches[1].data_ptr() - batches[0].data_ptr() if nb > 1 else 0
        return nb == 1 or (step > 0 and all(b.shape == batches[0].shape and b.stride() == batches[0].stride() and
                                            b.data_ptr() - batches[0].data_ptr() == k * step for k, b in enumerate(batches)))

    def planned(self, u: int, input: torch.Tensor, out: torch.Tensor = None) -> torch.Tensor:
        """The train-mode forward of update ``u`` of the last ``plan``: normalise with the statistics after that update."""
        x = input.detach()
        mean, var = self._plan_seq[u, 0], self._plan_seq[u, 1]
        if x.is_contiguous():
            y = out if out is not None else torch.empty_like(x)
            return ops.rms_normalize(x, mean, var, y, eps=self.epsilon)
        y = out if out is not None else torch.empty(x.shape[0] * x.shape[1], self.insize, dtype=torch.float32, device=x.device)
        ops.rms_normalize_slabs(x, mean, var, y, eps=self.epsilon)
        return y

    def planned_group(self, u0: int, batches, out: torch.Tensor = None) -> torch.Tensor:
        """The train-mode forwards of updates ``u0 .. u0 + len(batches) - 1`` of the last ``plan`` in ONE launch (the minibatches
        of one mini-epoch: equally spaced slab views of one rollout tensor).  Returns (len(batches), m, insize)."""
        nb, m = len(batches), batches[0].numel() // self.insize
        if u0 < 0 or u0 + nb > self._plan_seq.shape[0]:
            raise ValueError("u0 + len(batches) exceeds the planned updates")
        y = out if out is not None else torch.empty(nb, m, self.insize, dtype=torch.float32, device=batches[0].device)
        if not self._equally_spaced(batches):             # separately allocated minibatches: one launch each
            for k, b in enumerate(batches):
                self.planned(u0 + k, b, out=y[k].view(b.shape) if b.is_contiguous() else y[k])
            return y
        return ops.rms_normalize_slabs_batched([b.detach() for b in batches], self._plan_seq, u0, y, eps=self.epsilon)

    def forward(self, input: torch.Tensor, unnorm: bool = False, out: torch.Tensor = None) -> torch.Tensor:
        x = input.detach()
        if x.dtype != torch.float32:
            x = x.float()
        slab = (not x.is_contiguous()) and x.dim() == 3 and x.stride(2) == 1 and x.stride(1) == x.shape[2] \
            and x.shape[2] == self.insize
        if not slab:
            x = x.contiguous()
        if x.shape[-1] != self.insize and not (self.insize == 1 and x.dim() == 1):
            raise ValueError(f"expected last dim {self.insize}, got {tuple(x.shape)}")
        local = self.process_group is None or not bdist.is_distributed(self.process_group)
        if self.training and not unnorm:
            # train forward: moments (pivot = running_mean, in place) -> fold + snapshot -> [SUM all-reduce over ranks] -> merge +
            # normalise; ONE C call when local
            self._workspace(x.device)
            rows = x.shape[0] * x.shape[1] if slab else x.numel() // self.insize
            y = out if out is not None else (torch.empty(rows, self.insize, dtype=torch.float32, device=x.device) if slab
                                             else torch.empty_like(x))
            count = self.count.view(1)
            if local:
                ops.rms_train_forward(x, self.running_mean, self.running_var, count, y, self._scratch, eps=self.epsilon)
            else:
                ops.rms_moments_ext(x, self.running_mean, self.running_var, count, self._acc_ext, self._scratch)
                bdist.allreduce_sum_(self._acc_ext[:1 + 2 * self.insize], self.process_group)
                ops.rms_merge_normalize(x, self._acc_ext, self.running_mean, self.running_var, count, y, eps=self.epsilon)
            return y
        if self.training:
            self.update(x)
        if slab:
            # minibatch read in place from time-major storage: the output is the contiguous (T*E, insize) network input
            y = out if out is not None else torch.empty(x.shape[0] * x.shape[1], self.insize, dtype=torch.float32, device=x.device)
            ops.rms_normalize_slabs(x, self.running_mean, self.running_var, y, eps=self.epsilon, unnorm=unnorm)
            return y
        y = out if out is not None else torch.empty_like(x)
        ops.rms_normalize(x, self.running_mean, self.running_var, y, eps=self.epsilon, unnorm=unnorm)
        return y
