"""TEST INFRASTRUCTURE (oracle): CPU restatement of the batching done by the reference's decode scripts around
``batch_fast_generate`` - ``decode_generator`` / ``pad_list`` of src/bin/decode_cswnv_laplace-shift1.py:30-113 and
src/bin/decode_dswnv_softmax.py:54-117 - over in-memory features.  Only tests/ may import this module; the product
never does.
Parity: PINNED.  tests/golden_decode/*.npz are recordings of the reference scripts' own ``main()`` run end to end
(tests/golden/make_decode_golden.py: h5py / soundfile / torchvision stubbed, everything else unmodified);
tests/test_decode_driver.py holds both this restatement and the product's driver to them.  The mu-law maps are pinned
against vectors generated from the reference module (tests/golden_aux/mulaw.npz)."""
import math
import os

import numpy as np


def pad_list(batch_list):
    # decode_cswnv_laplace-shift1.py:30-48
    batch_size = len(batch_list)
    maxlen = max([batch.shape[0] for batch in batch_list])
    n_feats = batch_list[0].shape[-1]
    batch_pad = np.zeros((batch_size, maxlen, n_feats))
    for idx, batch in enumerate(batch_list):
        batch_pad[idx, :batch.shape[0]] = batch
    return batch_pad


def decode_generator(feats, feat_list, batch_size, seg, upsampling_factor, kind="cswnv"):
    """feats: {featfile: ndarray [T, C]}.  Yields (feat_ids, batch_x, batch_h [B, C, T_max], n_samples_list)."""
    # :77-84 sort with the feature length, divide into batch list
    shape_list = [feats[f].shape[0] for f in feat_list]
    idx = np.argsort(shape_list)
    feat_list = [feat_list[i] for i in idx]
    n_batch = math.ceil(len(feat_list) / batch_size)
    batch_lists = [f.tolist() for f in np.array_split(feat_list, n_batch)]
    for batch_list in batch_lists:
        batch_x, batch_h, n_samples_list, feat_ids = [], [], [], []
        for featfile in batch_list:
            # :93-99 (cswnv) seed zeros(seg); (dswnv :94-99) mu-law code of a zero sample
            if kind == "cswnv":
                x = np.zeros(seg)
            else:
                mu = 255
                x = np.floor((np.sign(np.zeros(1)) * np.log(1 + mu * 0.0) / np.log(1 + mu) + 1) / 2 * mu + 0.5).astype(np.int64)
            h = feats[featfile]
            batch_x += [x]
            batch_h += [h]
            n_samples_list += [h.shape[0] * upsampling_factor]
            feat_ids += [os.path.basename(featfile).replace(".h5", "")]
        batch_x = np.stack(batch_x, axis=0)
        batch_h = np.transpose(pad_list(batch_h), (0, 2, 1)).astype(np.float32)     # :104-108 FloatTensor(...).transpose(1, 2)
        yield feat_ids, batch_x, batch_h, n_samples_list
