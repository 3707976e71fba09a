"""Stand-in for torch_geometric (absent from this image; the reference pins it only as
`torch-geometric>=2.3.0`, requirements.txt:4). TEST INFRASTRUCTURE ONLY: it exists so that
`oracle/make_golden.py` can import the unmodified reference from /root/reference in the build
container and record golden vectors.  It restates exactly the PyG semantics the reference's hot
path relies on:

  * add_self_loops appends arange(N) loops AFTER the given edges and removes nothing;
  * degree(index, N) counts occurrences;
  * flow is source_to_target: x_j = x[edge_index[0]], aggregation at edge_index[1];
  * aggr='add' is an index_add_ over dim 0 (sequential in edge order on CPU);
  * softmax(src, index, num_nodes) = exp(src - segmax[index]) / (segsum[index] + 1e-16).

Nothing in the product package imports this.
"""
__version__ = "0.0-shim"
