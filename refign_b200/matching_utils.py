"""Host-side mirror of the reference's helpers/matching_utils.py for the functions on the hot path
(``warp``, the confidence estimate, mapping -> flow); same names and argument meaning, sm_100a
kernels underneath (refign_b200.ops), no host synchronisation."""
import torch

from . import ops



def warp(x, flo, padding_mode='zeros', return_mask=False):
    """helpers/matching_utils.py:11-49 on the sm_100a kernel (fp32, zeros padding, align_corners=True)."""
    return ops.warp(x, flo, padding_mode=padding_mode, return_mask=return_mask)


def estimate_probability_of_confidence_interval_of_mixture_density(uncert_output, R=1.0):
    """P_R = 1 - exp(-R^2 / (2 exp(log_var))) (helpers/matching_utils.py:52-57)."""
    return ops.estimate_probability_of_confidence_interval_of_mixture_density(uncert_output, R)


def unnormalise_and_convert_mapping_to_flow(map, output_channel_first=True):
    """Mapping in [-1, 1] -> flow in pixels: flow = (m + 1) * (size - 1) / 2 - grid
    (helpers/matching_utils.py:77-128)."""
    if map.dim() == 3:  # unbatched [2,H,W] or [H,W,2]
        cf = map.shape[0] == 2
        out = unnormalise_and_convert_mapping_to_flow((map if cf else map.permute(2, 0, 1)).unsqueeze(0))[0]
        return out if output_channel_first else out.permute(1, 2, 0)
    m = map if map.shape[1] == 2 else map.permute(0, 3, 1, 2)
    B, _, H, W = m.shape
    xx = torch.arange(W, device=m.device, dtype=m.dtype).view(1, 1, 1, W)
    yy = torch.arange(H, device=m.device, dtype=m.dtype).view(1, 1, H, 1)
    flow = torch.cat(((m[:, 0:1] + 1) * (W - 1) / 2.0 - xx, (m[:, 1:2] + 1) * (H - 1) / 2.0 - yy), 1)
    return flow if output_channel_first else flow.permute(0, 2, 3, 1)
