"""spconv.pytorch.functional: indice_conv / indice_subm_conv / indice_inverse_conv on explicit pair
lists (reference layers/blocks_3d/det3d/centernet_backbone.py:253-276).  `filters` uses the layout of
that hand-rolled module, (kz, ky, kx, Cin, Cout) (centernet_backbone.py:168)."""
import torch

from unidistill_b200 import sparse as _sp


def _tables_from_pairs(indice_pairs, n_in, n_out):
    """Pair list (2,K,N) -> (nbr_out (n_out,K), nbr_in (n_in,K)); index bookkeeping only."""
    K = indice_pairs.shape[1]
    dev = indice_pairs.device
    valid = indice_pairs[0] >= 0
    ks = torch.arange(K, device=dev).view(K, 1).expand_as(indice_pairs[0])
    i_idx = indice_pairs[0][valid].long()
    o_idx = indice_pairs[1][valid].long()
    k_idx = ks[valid]
    nbr_out = torch.full((n_out, K), -1, dtype=torch.int32, device=dev)
    nbr_in = torch.full((n_in, K), -1, dtype=torch.int32, device=dev)
    nbr_out[o_idx, k_idx] = i_idx.int()
    nbr_in[i_idx, k_idx] = o_idx.int()
    return nbr_out, nbr_in


def _rulebook_from_pairs(features, filters, indice_pairs, num_activate_out, subm):
    ksize = [int(s) for s in filters.shape[:3]]
    n_in = int(features.shape[0])
    n_out = int(num_activate_out)
    nbr_out, nbr_in = _tables_from_pairs(indice_pairs, n_in if not subm else n_out, n_out)
    return _sp.Rulebook(subm, ksize, [1, 1, 1], [0, 0, 0], [1, 1, 1], [0, 0, 0], [0, 0, 0], n_in, n_out,
                        nbr_out, None if subm else nbr_in, None, 0)


def _to_param_layout(filters):
    # (kz,ky,kx,Cin,Cout) -> (Cout,kz,ky,kx,Cin)
    return filters.permute(4, 0, 1, 2, 3)


def indice_conv(features, filters, indice_pairs, indice_pair_num, num_activate_out, algo=None):
    rb = _rulebook_from_pairs(features, filters, indice_pairs, num_activate_out, subm=False)
    return _sp.sparse_conv(features, _to_param_layout(filters), None, rb, False)


def indice_subm_conv(features, filters, indice_pairs, indice_pair_num, num_activate_out, algo=None):
    rb = _rulebook_from_pairs(features, filters, indice_pairs, num_activate_out, subm=True)
    return _sp.sparse_conv(features, _to_param_layout(filters), None, rb, False)


def indice_inverse_conv(features, filters, indice_pairs, indice_pair_num, num_activate_out, algo=None):
    # features live on the *output* rows of the paired conv; result has num_activate_out rows
    ksize = [int(s) for s in filters.shape[:3]]
    n_small = int(features.shape[0])
    n_big = int(num_activate_out)
    nbr_out, nbr_in = _tables_from_pairs(indice_pairs, n_big, n_small)
    rb = _sp.Rulebook(False, ksize, [1, 1, 1], [0, 0, 0], [1, 1, 1], [0, 0, 0], [0, 0, 0], n_big, n_small,
                      nbr_out, nbr_in, None, 0)
    return _sp.sparse_conv(features, _to_param_layout(filters), None, rb, True)
