"""Stand-in for torchprofile (absent here).  Like the real profile_macs it jit-TRACES the model on the given inputs
(torchprofile/utils/trace.py: torch.jit._get_trace_graph) and walks the graph — the hazard SURVEY §7 names for a drop-in module — and returns a
MAC count: here only aten::linear / addmm / matmul / conv nodes that carry static shapes are counted, unknown nodes count 0 (the real package
warns "No handlers found" for them)."""
import torch

CALLS = []


def profile_macs(model, args=(), kwargs=None, reduction=sum):
    if not isinstance(args, (tuple, list)):
        args = (args,)
    graph, _ = torch.jit._get_trace_graph(model, tuple(args), kwargs)
    kinds = {}
    for node in graph.nodes():
        kinds[node.kind()] = kinds.get(node.kind(), 0) + 1
    CALLS.append(kinds)
    return sum(kinds.values())
