"""ORACLE (test infrastructure): a drop-in for popnet_b200's CUDA backend that runs the C oracle on the
host.  Tests inject it (``popnet_b200.evaluate._backend = OracleBackend()``) to check the host-side
logic (list packing, AP tail, re-materialisation of the reference's return types) without a GPU.
The product never imports this module.
"""
from . import c_oracle


class OracleBackend:
    name = "oracle-c"

    def pck(self, arrs, *, dist_th, iou_th, K):
        return c_oracle.eval_pck(arrs, dist_th=dist_th, iou_th=iou_th, K=K)

    def map_assign(self, arrs, *, thresh, K, D):
        return c_oracle.eval_map_assign(arrs, thresh=thresh, K=K, D=D)

    def decode(self, heat, paf, depth, params):
        return c_oracle.decode(heat, paf, depth, params)
