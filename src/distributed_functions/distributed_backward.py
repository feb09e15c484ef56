"""``backward(variables, grad_variables, retain_graph)`` under the reference's module path
(``/root/reference/src/distributed_functions/distributed_backward.py:38-90``: a vendored copy of torch-0.2's ``torch.autograd.backward`` that
``single_machine.py`` imported and never called).

The hook the authors wanted from it — "do something as each layer's gradient appears" — is what
``ps_pytorch_b200.models.layerwise.LayerwiseBackwardMixin`` provides (per-parameter callbacks fired by autograd while the rest of the
backward pass is still running).  This function keeps the old call signature for scripts that import it and adds the optional callback.
"""
import _path  # noqa: F401
import torch


def backward(variables, grad_variables=None, retain_graph=None, create_graph=None, retain_variables=None, on_grad=None):
    """Sum of gradients of ``variables`` w.r.t. the graph leaves.  ``retain_variables`` is the torch-0.2 spelling of ``retain_graph``.
    ``on_grad(param, grad)``, when given, is invoked for every leaf gradient the moment autograd produces it."""
    if retain_variables is not None and retain_graph is None:
        retain_graph = retain_variables
    variables = (variables,) if torch.is_tensor(variables) else tuple(variables)
    if grad_variables is not None:
        grad_variables = (grad_variables,) if torch.is_tensor(grad_variables) else tuple(grad_variables)
    handles = []
    if on_grad is not None:
        seen, stack = set(), [v.grad_fn for v in variables if v.grad_fn is not None]
        while stack:                                   # walk the graph once to find the AccumulateGrad leaves
            fn = stack.pop()
            if fn is None or fn in seen:
                continue
            seen.add(fn)
            if hasattr(fn, "variable"):
                leaf = fn.variable
                handles.append(leaf.register_hook(lambda g, leaf=leaf: on_grad(leaf, g)))
            stack.extend(nf for nf, _ in fn.next_functions)
    try:
        torch.autograd.backward(variables, grad_variables, retain_graph=retain_graph, create_graph=bool(create_graph))
    finally:
        for h in handles:
            h.remove()
