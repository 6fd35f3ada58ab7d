"""Whole-step CUDA graph for MHCT training (the step-loop overheads that cap multi-GPU scaling).

With the Splat / Slice path down to ~10 % of a ScanObjectNN training step, the step is a stream of ~4 000 small kernel
launches (~7 000 with SyncBatchNorm, whose forward / backward are Python autograd functions of a dozen tiny kernels
and one small NCCL collective each: 106 layers in model_zoo/scanobject/classifier.py).  On one GPU the host keeps up;
under DDP + SyncBN it does not, and the step time is set by Python / launch overhead instead of by the GPU
(profiles/r02_train_prof_*.txt).  Capturing forward + backward + optimizer step -- including the NCCL collectives of
SyncBN and of DDP's gradient buckets, which are graph-capturable -- into ONE CUDA graph removes that overhead: a step
becomes a single graph launch per rank.  The kernels of libctb200 only enqueue work on the stream they are given and
never synchronise, so they capture like any other kernel.

    step = GraphedTrainStep(model, optimizer, loss_fn, example_inputs)      # model may be DDP(SyncBN(model))
    outputs = step(*inputs)       # copies the inputs into the static buffers, replays, returns the static outputs

Rules inherited from CUDA graphs: fixed shapes, no host synchronisation inside the step, optimizer created with
capturable=True.  Under DDP set TORCH_NCCL_ASYNC_ERROR_HANDLING=0 before init_process_group and construct DDP on the
side stream used for the warm-up (PyTorch's documented recipe); eleven eager iterations precede the capture.
"""
import torch


class GraphedTrainStep:
    def __init__(self, model, optimizer, loss_fn, example_inputs, warmup=11):
        """loss_fn(model_outputs, *example_inputs[n_model_inputs:]) -> (loss, extras); the first tensor of
        example_inputs feeds the model, the rest go to loss_fn."""
        self.model, self.optimizer, self.loss_fn = model, optimizer, loss_fn
        self.static_in = [t.clone() for t in example_inputs]
        side = torch.cuda.Stream(device=self.static_in[0].device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._eager_step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        optimizer.zero_grad(set_to_none=True)
        with torch.cuda.graph(self.graph):
            outs = model(self.static_in[0])
            loss, extras = loss_fn(outs, *self.static_in[1:])
            loss.backward()
            optimizer.step()
        self.static_out = (outs, loss, extras)

    def _eager_step(self):
        self.optimizer.zero_grad(set_to_none=True)
        outs = self.model(self.static_in[0])
        loss, _ = self.loss_fn(outs, *self.static_in[1:])
        loss.backward()
        self.optimizer.step()

    def __call__(self, *inputs):
        for dst, src in zip(self.static_in, inputs):
            dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.static_out
