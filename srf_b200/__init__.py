"""srf_b200: B200-native (sm_100a) capsule-routing hot path of sephiroce/srf.

Public surface:
    SequenceRouter   -- drop-in for tfsr.model.sequence_router_naive.SequenceRouter
    RoutingStack     -- the routing stack alone (primary capsules -> CTC logits)
    routing          -- tensor-level wrappers of the C-ABI (include/srf_b200.h)
    autograd         -- the stack's forward/backward pair as one differentiable torch op
    checkpoint       -- weights interchange with the reference layouts, checkpoints, averaging
"""
from . import _lib, autograd, checkpoint, routing, training  # noqa: F401
from .sequence_router import HostPipeline, RoutingStack, SequenceRouter, layer_shapes  # noqa: F401

__all__ = ["routing", "training", "checkpoint", "autograd", "RoutingStack", "SequenceRouter", "HostPipeline", "layer_shapes"]
