import inspect
import torch


class MessagePassing(torch.nn.Module):
    """propagate() = gather `_j` arguments at edge_index[0], call message(), sum at edge_index[1]."""

    def __init__(self, aggr="add", flow="source_to_target", node_dim=0, **kwargs):
        super().__init__()
        assert aggr == "add" and flow == "source_to_target" and node_dim == 0
        self.aggr = aggr

    def propagate(self, edge_index, size=None, **kwargs):
        src, dst = edge_index[0], edge_index[1]
        params = inspect.signature(self.message).parameters
        args = {}
        n_out = None
        for name in params:
            if name.endswith("_j"):
                base = kwargs[name[:-2]]
                n_out = base.size(0)
                args[name] = base.index_select(0, src)
            elif name.endswith("_i"):
                base = kwargs[name[:-2]]
                n_out = base.size(0)
                args[name] = base.index_select(0, dst)
            else:
                args[name] = kwargs.get(name)
        msg = self.message(**args)
        if size is not None and size[1] is not None:
            n_out = size[1]
        out = torch.zeros((n_out,) + tuple(msg.shape[1:]), dtype=msg.dtype, device=msg.device)
        return out.index_add_(0, dst, msg)
