import copy
import torch


class Data:
    """Attribute bag with PyG>=2.3 behaviour for unset `batch`/`pos`/...: they read as None."""

    _KNOWN = ("x", "edge_index", "edge_attr", "y", "pos", "batch")

    def __init__(self, x=None, edge_index=None, edge_attr=None, y=None, pos=None, **kwargs):
        self.x, self.edge_index, self.edge_attr, self.y, self.pos = x, edge_index, edge_attr, y, pos
        self.batch = None
        for k, v in kwargs.items():
            setattr(self, k, v)

    @property
    def num_nodes(self):
        if "_num_nodes" in self.__dict__ and self.__dict__["_num_nodes"] is not None:
            return self.__dict__["_num_nodes"]
        return self.x.size(0) if self.x is not None else 0

    @num_nodes.setter
    def num_nodes(self, v):
        self.__dict__["_num_nodes"] = v

    @property
    def num_edges(self):
        return self.edge_index.size(1) if self.edge_index is not None else 0

    def clone(self):
        out = copy.copy(self)
        for k, v in self.__dict__.items():
            if torch.is_tensor(v):
                out.__dict__[k] = v.clone()
        return out

    def to(self, device):
        out = copy.copy(self)
        for k, v in self.__dict__.items():
            if torch.is_tensor(v):
                out.__dict__[k] = v.to(device)
        return out


class Batch(Data):
    @classmethod
    def from_data_list(cls, data_list):
        xs, eis, eas, poss, batch = [], [], [], [], []
        off = 0
        for i, d in enumerate(data_list):
            n = d.x.size(0)
            xs.append(d.x)
            eis.append(d.edge_index + off)
            if d.edge_attr is not None:
                eas.append(d.edge_attr)
            if d.pos is not None:
                poss.append(d.pos)
            batch.append(torch.full((n,), i, dtype=torch.long))
            off += n
        out = cls(x=torch.cat(xs), edge_index=torch.cat(eis, dim=1),
                  edge_attr=torch.cat(eas) if eas else None,
                  pos=torch.cat(poss) if poss else None)
        out.batch = torch.cat(batch)
        out.num_graphs = len(data_list)
        return out
