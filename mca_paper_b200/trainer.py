"""Fused training step: H2D of the batch -> forward -> all-pairs loss -> backward -> (DP all-reduce) -> clip + AdamW
-> weight re-pack, replayed from CUDA graphs (shapes are static: Q6, the batch size is baked into the model).

Replaces the reference hot loop train_accel_gpu.py:110-119 (move_to, model(batch), zero_grad, backward,
clip_grad_norm_, optimizer.step, lr_scheduler.step) without its per-step host synchronisations
(train_accel_gpu.py:126-130 log every scalar with .to("cpu"); here the loss stays on the device until asked for).

Data parallelism: one process per GPU, identical replicas, B samples per rank.  The exchange of the pooled
block [B,R,512], the reduce-scatter of its gradient and the gradient reduction + parameter exchange of the optimiser all
run as our own push / pull kernels over P2P-mapped symmetric memory behind flag barriers (loss.cu, optim.cu), so the whole
step is ONE captured graph at any world size.  MCA_P2P=0 selects the NCCL form (all_gather, reduce_scatter, all_reduce
between graph segments).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from .engine import Engine


class Trainer:
    def __init__(self, model, lr: float = 1e-4, clip: float = 2.0, weight_decay: float = 0.01, betas=(0.9, 0.999),
                 eps: float = 1e-8, schedule: str = "constant", warmup_steps: int = 0, total_steps: int = 1,
                 use_graphs: bool = True, process_group=None, scheduler_stride: Optional[int] = None):
        self.model = model
        self.eng: Engine = model.engine
        self.eng.ensure_flat()
        dist = torch.distributed.is_available() and torch.distributed.is_initialized()
        if scheduler_stride is None:  # accelerate's scheduler wrapper steps once per process (train_accel_gpu.py:93,119)
            scheduler_stride = torch.distributed.get_world_size(process_group) if dist else 1
        self.eng.configure_optimizer(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, max_norm=clip,
                                     schedule=schedule, warmup_steps=warmup_steps, total_steps=total_steps,
                                     scheduler_stride=scheduler_stride)
        if dist:
            self.eng.set_distributed(torch.distributed.get_world_size(process_group),
                                     torch.distributed.get_rank(process_group), process_group)
        self.use_graphs = use_graphs
        self._graphs = None
        self._dev_batch: Optional[Dict[str, Dict[str, torch.Tensor]]] = None
        self._pinned: Optional[Dict[str, Dict[str, torch.Tensor]]] = None
        self._pinned_sets = None
        self.eng.pack_weights()
        self.h2d_bytes = 0
        self.kernel_launches_per_step = None

    # ------------------------------------------------------------------------------------------ staging
    # The host fills a PINNED buffer and enqueues an asynchronous H2D copy from it; the step never synchronises, so the
    # host can be several steps ahead of the device.  A pinned buffer may therefore only be rewritten once the copy that
    # reads it has finished: there are two pinned sets, used alternately, each guarded by a CUDA event recorded right
    # after its H2D copies (the host waits on that event, not on the device as a whole, before refilling the set).
    def _ensure_staging(self, host_batch):
        if self._dev_batch is not None:
            return
        dev = self.eng.device
        self._dev_batch, self._pinned_sets = {}, [{}, {}]
        n = 0
        for m, d in host_batch.items():
            self._dev_batch[m] = {}
            for ps in self._pinned_sets:
                ps[m] = {}
            for k, v in d.items():
                self._dev_batch[m][k] = torch.empty(v.shape, dtype=v.dtype, device=dev)
                for ps in self._pinned_sets:
                    ps[m][k] = torch.empty(v.shape, dtype=v.dtype, pin_memory=True)
                n += v.numel() * v.element_size()
        self._pinned = self._pinned_sets[0]           # the set filled last (h2d() / prefetch() read it)
        self._pin_last = 0
        self._pin_free = [None, None]                 # event: the copies out of pinned set i have completed
        self._pin_next = 0
        self.h2d_bytes = n

    def _fill_pinned(self, host_batch):
        """Host batch -> the next pinned set (after its previous H2D copies have drained).  Returns the set index."""
        self._ensure_staging(host_batch)
        i = self._pin_next
        self._pin_next ^= 1
        if self._pin_free[i] is not None:
            self._pin_free[i].synchronize()
        for m, d in host_batch.items():
            for k, v in d.items():
                self._pinned_sets[i][m][k].copy_(v)
        self._pinned = self._pinned_sets[i]
        self._pin_last = i
        return i

    def _mark_pinned_read(self, i, stream=None):
        ev = self._pin_free[i]
        if ev is None:
            ev = self._pin_free[i] = torch.cuda.Event()
        ev.record(stream if stream is not None else torch.cuda.current_stream())

    def stage(self, host_batch):
        """Copy a host batch into a pinned staging set and enqueue its H2D copy on the current stream."""
        i = self._fill_pinned(host_batch)
        for m, d in self._pinned_sets[i].items():
            for k, v in d.items():
                self._dev_batch[m][k].copy_(v, non_blocking=True)
        self._mark_pinned_read(i)

    def h2d(self):
        """Enqueue only the H2D copies from the pinned set filled last (the caller keeps it unchanged meanwhile)."""
        for m, d in self._pinned.items():
            for k, v in d.items():
                self._dev_batch[m][k].copy_(v, non_blocking=True)

    def stage_device(self, dev_batch):
        """Hand a batch that is already ON THE DEVICE (e.g. the output of collate.DeviceCollator) to the step's input
        buffers: device-to-device copies on the current stream, no host round trip."""
        self._ensure_staging(dev_batch)
        for m, d in dev_batch.items():
            for k, v in d.items():
                self._dev_batch[m][k].copy_(v, non_blocking=True)

    def step_device(self, dev_batch):
        """One optimisation step on a device-resident batch (DeviceCollator output)."""
        self.stage_device(dev_batch)
        return self.step_staged()

    # ---- pipelined input path: the H2D copy of the next batch runs on a copy stream under the current step
    def _ensure_pipeline(self):
        if getattr(self, "_copy_stream", None) is not None:
            return
        self._copy_stream = torch.cuda.Stream()
        self._slots = [{m: {k: torch.empty_like(v) for k, v in d.items()} for m, d in self._dev_batch.items()}
                       for _ in range(2)]
        self._ready = [torch.cuda.Event() for _ in range(2)]      # H2D into the slot finished
        self._consumed = [torch.cuda.Event() for _ in range(2)]   # the step has copied the slot into its inputs
        for e in self._consumed:
            e.record()

    def prefetch(self, slot: int, host_batch=None):
        """Start the H2D copy of a host batch into device staging slot `slot` on the copy stream.  With `host_batch` the
        batch is first copied into a free pinned set; without it the pinned set filled last is sent again (the
        benchmark's fixed synthetic batch)."""
        if host_batch is not None:
            i = self._fill_pinned(host_batch)
        else:
            i = self._pin_last
        self._ensure_pipeline()
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(self._consumed[slot])
            for m, d in self._pinned_sets[i].items():
                for k, v in d.items():
                    self._slots[slot][m][k].copy_(v, non_blocking=True)
            self._ready[slot].record()
            self._mark_pinned_read(i, self._copy_stream)

    def step_from_slot(self, slot: int):
        """Run one step on the batch prefetched into `slot` (device-to-device hand-over into the graph's inputs)."""
        cur = torch.cuda.current_stream()
        cur.wait_event(self._ready[slot])
        for m, d in self._slots[slot].items():
            for k, v in d.items():
                self._dev_batch[m][k].copy_(v, non_blocking=True)
        self._consumed[slot].record()
        return self.step_staged()

    # ------------------------------------------------------------------------------------------ step pieces
    def _seg_forward(self):
        self._pooled = self.eng.trunk_forward(self._dev_batch)

    def _seg_loss(self):
        eng = self.eng
        eng.loss_forward(self._pooled)
        eng.flat_grad.zero_()
        self._dpooled = eng.loss_backward(eng.ws["w_default"])

    def _seg_backward(self):
        self.eng.trunk_backward(self._dpooled)

    def _seg_optim(self):
        self.eng.optimizer_step()

    def _run_eager(self):
        self._seg_forward()
        self._seg_loss()
        self._seg_backward()
        self._seg_optim()

    def _capture(self):
        eng = self.eng
        # warm-up on a side stream (allocations, attribute setting, NCCL init) before capture; the two warm-up steps
        # are real optimiser steps, so the training state is put back afterwards: N calls of step() == N updates.
        # (Peer-memory mode: every rank's last action in a step is the barrier that publishes the parameter stores, so
        # after the local synchronize nobody writes into this rank's buffers any more.)
        snap = eng.snapshot_state()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):
                self._run_eager()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        eng.restore_state(snap)
        torch.cuda.synchronize()
        if eng.world == 1 or eng._p2p is not None:
            # single GPU, or data parallel over peer memory (every exchange is one of our kernels): ONE graph
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._run_eager()
            self._graphs = [("graph", g)]
        else:
            # collectives stay outside the graphs: [fwd] all_gather+loss(eager: contains 2 collectives) [bwd] all_reduce+optim
            g1, g3 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(g1):
                self._seg_forward()
            with torch.cuda.graph(g3, pool=g1.pool()):   # consumes the buffers the (eager) loss segment fills
                self._seg_backward()
            self._graphs = [("graph", g1), ("eager", self._seg_loss), ("graph", g3), ("eager", self._seg_optim)]
        torch.cuda.synchronize()

    def step_staged(self):
        """Run one optimisation step on the batch currently in the device staging buffers; returns the device
        tensor [loss, fcl_loss, no-fcl_loss, #valid losses] (no synchronisation)."""
        if not self.use_graphs:
            self._run_eager()
        else:
            if self._graphs is None:
                self._capture()
            for kind, g in self._graphs:
                if kind == "graph":
                    g.replay()
                else:
                    g()
        self.eng.check_p2p()  # host read of a pinned flag: a peer that missed a barrier fails the run, not just the number
        return self.eng.ws["summary"]

    def optimizer_state_dict(self):
        """torch.optim.AdamW-layout optimiser state (gathered across ranks under peer-memory data parallelism)."""
        return self.eng.optimizer_state_dict()

    def load_optimizer_state_dict(self, sd):
        self.eng.load_optimizer_state_dict(sd)

    def step(self, host_batch):
        """End-to-end step from a HOST batch (dict of dict of CPU tensors): pinned staging, H2D, fused step."""
        self.stage(host_batch)
        return self.step_staged()
