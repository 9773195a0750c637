"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the query selection's top-k.

The reference selects the decoder's queries with
    topk_ind = torch.topk(enc_outputs_scores.max(-1).values, self.num_queries, dim=1).indices
(ultralytics/nn/modules/head.py:1240 in ManbaWorldDecoder._get_decoder_input, :437 in RTDETRDecoder): per image, the
indices of the k highest-scoring tokens, best first.  torch.topk leaves the order of EQUAL scores unspecified; the kernel
under test (tamtr_b200/csrc/topk.cu) documents one -- lower index first -- so the oracle is the stable descending sort,
whose first k entries are a valid torch.topk result for every input and THE torch.topk result wherever the scores are
distinct.  Pinned on the CPU against torch.topk itself (tests/test_oracle.py): equal indices on distinct scores, equal
values always.  NaN ranks first and -0.0 == +0.0, as for torch.
"""
import torch


def topk_indices(scores, k):
    """scores [B, n] (any float dtype, any device) -> int64 [B, k] on the CPU: (score descending, index ascending)."""
    return torch.sort(scores.detach().cpu(), dim=1, descending=True, stable=True).indices[:, :k].contiguous()
