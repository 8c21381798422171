"""ocbo-b200: the GP posterior / Thompson-sampling hot path of fusion-ml/OCBO on B200 (sm_100a)."""
