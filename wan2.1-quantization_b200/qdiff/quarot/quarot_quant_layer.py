"""QuaRot layer (ViDiT-Q/quant_utils/qdiff/quarot/quarot_quant_layer.py:7-70), integer execution:

    rotation_matrix R = diag(+-1) . H / sqrt(n)  (random_hadamard_matrix)               (:27-28)
    weights : (W.double() @ R).float() re-quantised per out-channel (offline)            (:30-45)
    forward : (x.double() @ R) -> per-token quantizer -> tcgen05 int8 GEMM              (:47-70)

The activation rotation runs as a Hadamard transform (butterflies + one order-K block, qdiff.quarot.quarot_utils),
in fp32 on the device instead of the reference's dense fp64 matmul; a fused in-shared-memory transform inside the
quantizer kernel is the remaining step of SURVEY §8 f-2."""
import torch

from qdiff.base.quant_layer import QuantizedLinear
from qdiff.base.quant_layer import ActPlan
from qdiff.quarot.quarot_utils import hadamard_kernel_plan, matmul_hadU, random_hadamard_matrix


class RotationMixin:
    """x @ R.  A matrix produced by this package's random_hadamard_matrix is R = diag(s) . Hn / sqrt(n): multiply by the
    sign vector, then the structured transform (butterflies + one order-K block).  Any other orthogonal matrix (e.g. one
    a reference checkpoint was calibrated with) is applied as a dense fp32 product."""

    def get_rotation_matrix(self):
        dev = self.fp_module.weight.device
        self.rotation_matrix = random_hadamard_matrix(self.in_features, dev)
        self._rot_plan = None

    def _plan_rotation(self):
        R = self.rotation_matrix
        if R is None:
            raise RuntimeError("rotation_matrix not set (run PTQ or load_quant_param_dict first)")
        n = R.shape[0]
        Hn = matmul_hadU(torch.eye(n, dtype=torch.float64, device=R.device))
        s = torch.sign((R.double() * Hn).sum(dim=1))
        if torch.allclose(s.view(-1, 1) * Hn, R.double(), atol=1e-9):
            return ("structured", s.float(), R)
        return ("dense", R.float(), R)

    def rotation_sign(self):
        """s of R = diag(s) . H_n / sqrt(n) (float32 [n]) for a structured rotation, else None."""
        plan = getattr(self, "_rot_plan", None)
        if plan is None or plan[2] is not self.rotation_matrix:
            plan = self._rot_plan = self._plan_rotation()
        return plan[1] if plan[0] == "structured" else None

    def _rotation_act_plan(self, device, mask=None):
        """ActPlan of x -> (x * mask) @ R for the fused kernel, or None (dense rotation / unsupported size)."""
        cached = getattr(self, "_act_plan_cache", None)
        if cached is not None and cached[0] is self.rotation_matrix and cached[1] is mask and cached[2] == device:
            return cached[3]
        plan = None
        s = self.rotation_sign()
        if s is not None:
            plan = ActPlan.rotation(self.in_features, s, mask, device)
        self._act_plan_cache = (self.rotation_matrix, mask, device, plan)
        return plan

    def _rotate(self, x2d):
        plan = getattr(self, "_rot_plan", None)
        if plan is None or plan[2] is not self.rotation_matrix:
            plan = self._rot_plan = self._plan_rotation()
        kind, t, _ = plan
        if kind == "structured":
            return matmul_hadU(x2d.float() * t.to(x2d.device).view(1, -1)).to(x2d.dtype)
        return (x2d.float() @ t.to(x2d.device)).to(x2d.dtype)


class QuarotQuantizedLinear(RotationMixin, QuantizedLinear):
    def __init__(self, in_features, out_features, bias, device, quant_config, fp_module):
        super().__init__(in_features, out_features, bias, device, quant_config, fp_module)
        self.rotation_matrix = None
        self._rot_plan = None
        self._rotated_weight = None

    def update_quantized_weight_rotated(self):
        self.w_quantizer.init_done = False
        R = self.rotation_matrix.to(self.fp_module.weight.device)
        self._rotated_weight = torch.matmul(self.fp_module.weight.data.double(), R).float()
        self.weight.data = self.w_quantizer(self._rotated_weight)
        self.w_quantizer.init_done = True
        self.invalidate_int_weight()

    def _weight_for_codes(self):
        return self._rotated_weight if self._rotated_weight is not None else self.fp_module.weight

    def _prepare_activation(self, x2d):
        return self._rotate(x2d)

    def _act_plan(self, device):
        return self._rotation_act_plan(device)
