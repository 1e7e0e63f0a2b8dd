"""Token-stream batching with the reference's semantics (commu/model/dataset.py:18-237): the four
object-array .npy files, pad(0) prepended as start token, per-column stream packing with reset
flags for training, contiguous sample groups for evaluation.  Written around per-column cursors and
pinned staging buffers; `BaseVocab` stays importable here because checkpoints pickle it."""
import os

import numpy as np
import torch

from commu.preprocessor.encoder.event_tokens import TOKEN_OFFSET


class BaseVocab:
    def __init__(self):
        self.vec_len = 0

    @property
    def pad_id(self):
        return 0

    def __len__(self):
        return TOKEN_OFFSET.VOCAB_SIZE.value


def _read_split(data_dir, tag):
    inp = np.load(os.path.join(data_dir, "input_%s.npy" % tag), allow_pickle=True)
    tgt = np.load(os.path.join(data_dir, "target_%s.npy" % tag), allow_pickle=True)
    return [np.concatenate((np.asarray(a, dtype=int), b)) for a, b in zip(inp, tgt)]


class ComMUDataset:
    def __init__(self, data_dir, cfg=None, verbose=True):
        self._vocab = BaseVocab()
        self.cfg = cfg
        pad = self._vocab.pad_id
        start = lambda seqs: [torch.from_numpy(np.concatenate(([pad], s)).astype(np.int64)) for s in seqs]
        self._splits = {"train": start(_read_split(data_dir, "train"))}
        val = start(_read_split(data_dir, "val"))
        self._splits["valid"] = val          # the reference loads the val files for both (dataset.py:80-86)
        self._splits["test"] = val
        self._lengths = {k: np.array([len(s) for s in v], dtype=np.int32) for k, v in self._splits.items()}
        if verbose:
            print("Loaded Data, #Samples Train/Val/Test:{}/{}/{}".format(
                *(len(self._splits[k]) for k in ("train", "valid", "test"))))

    vocab = property(lambda self: self._vocab)
    train_data = property(lambda self: self._splits["train"])
    valid_data = property(lambda self: self._splits["valid"])
    test_data = property(lambda self: self._splits["test"])
    train_seq_length = property(lambda self: self._lengths["train"])
    valid_seq_length = property(lambda self: self._lengths["valid"])
    test_seq_length = property(lambda self: self._lengths["test"])

    def get_iterator(self, batch_size, bptt, device, split="train", do_shuffle=True, seed=None):
        if split not in self._splits:
            raise NotImplementedError(split)
        seqs, lens = self._splits[split], self._lengths[split]
        n = len(seqs)
        pad = self._vocab.pad_id
        pin = torch.cuda.is_available() and torch.device(device).type == "cuda"

        def gen():
            """Batches are produced ONE AHEAD of the consumer: batch k+1 is packed into the next pinned staging set
            and its host->device copy is issued on a side stream before batch k is handed out, so packing and the
            copy overlap the training step of batch k.  A staging set is reused only after the
            copy that read it has completed (one CUDA event per set) - there is no per-batch stream synchronise."""
            assert batch_size < n
            order = np.arange(n)
            rng = np.random.RandomState(seed) if do_shuffle else None
            if do_shuffle:
                rng.shuffle(order)
            st = dict(cur_sample=list(range(batch_size)), cur_pos=[0] * batch_size, upcoming=batch_size, k=0)
            depth = 3
            stage = [(torch.empty(bptt, batch_size, dtype=torch.int64, pin_memory=pin),
                      torch.empty(bptt, batch_size, dtype=torch.int64, pin_memory=pin),
                      torch.empty(batch_size, dtype=torch.bool, pin_memory=pin)) for _ in range(depth if pin else 1)]
            copied = [None] * len(stage)
            copy_stream = torch.cuda.Stream(device) if pin else None

            def produce():
                slot = st["k"] % len(stage)
                st["k"] += 1
                data, target, reset = stage[slot]
                if copied[slot] is not None:
                    copied[slot].synchronize()           # the copy issued `depth` batches ago has read this set
                while True:
                    data.fill_(pad)
                    target.fill_(pad)
                    reset.fill_(False)
                    n_tok = 0
                    cur_sample, cur_pos = st["cur_sample"], st["cur_pos"]
                    for col in range(batch_size):
                        # a column whose sample is exhausted moves to the next unclaimed sample
                        while cur_sample[col] < n and cur_pos[col] + 1 >= lens[order[cur_sample[col]]]:
                            cur_sample[col], cur_pos[col] = st["upcoming"], 0
                            st["upcoming"] += 1
                            reset[col] = True
                        if cur_sample[col] >= n:
                            continue
                        seq = seqs[order[cur_sample[col]]]
                        p = cur_pos[col]
                        take = min(len(seq) - 1 - p, bptt)
                        data[:take, col] = seq[p:p + take]
                        target[:take, col] = seq[p + 1:p + 1 + take]
                        cur_pos[col] = p + take
                        n_tok += take
                    if n_tok > 0:
                        break
                    if not do_shuffle:
                        return None
                    rng.shuffle(order)
                    st["cur_sample"], st["cur_pos"], st["upcoming"] = list(range(batch_size)), [0] * batch_size, batch_size
                if not pin:      # (a CPU "copy" would alias the staging set that the look-ahead overwrites)
                    return (data.to(device, copy=True), target.to(device, copy=True), reset.to(device, copy=True), n_tok)
                with torch.cuda.stream(copy_stream):
                    out = (data.to(device, non_blocking=True), target.to(device, non_blocking=True),
                           reset.to(device, non_blocking=True))
                    ev = torch.cuda.Event()
                    ev.record(copy_stream)
                copied[slot] = ev
                return out + (n_tok, ev)

            nxt = produce()
            while nxt is not None:
                cur = nxt
                nxt = produce()                          # pack + copy the next batch before this one is consumed
                if pin:
                    d_, t_, r_, n_tok, ev = cur
                    consumer = torch.cuda.current_stream(device)
                    consumer.wait_event(ev)
                    for x in (d_, t_, r_):
                        x.record_stream(consumer)
                    yield d_, t_, r_, n_tok
                else:
                    yield cur

        return gen

    def eval_iterator(self, batch_size, bptt, device, split="valid", local_rank=0, world_size=0):
        if split not in ("valid", "test"):
            raise NotImplementedError(split)
        seqs, lens = self._splits[split], self._lengths[split]
        if world_size > 0:
            per = len(seqs) // world_size
            lo = per * local_rank
            hi = len(seqs) if local_rank == world_size - 1 else per * (local_rank + 1)
            seqs, lens = seqs[lo:hi], lens[lo:hi]
        pad = self._vocab.pad_id

        def gen():
            for g0 in range(0, len(seqs), batch_size):
                group = range(g0, min(g0 + batch_size, len(seqs)))
                longest = max(lens[i] for i in group)
                first = True
                for begin in range(0, longest - 1, bptt):
                    data = torch.full((bptt, batch_size), pad, dtype=torch.int64)
                    target = torch.full((bptt, batch_size), pad, dtype=torch.int64)
                    n_tok = 0
                    for i in group:
                        if lens[i] > begin + 1:
                            take = min(begin + bptt, lens[i] - 1) - begin
                            data[:take, i - g0] = seqs[i][begin:begin + take]
                            target[:take, i - g0] = seqs[i][begin + 1:begin + 1 + take]
                            n_tok += take
                    yield data.to(device), target.to(device), first, n_tok
                    first = False

        return gen


def write_synthetic_dataset(data_dir, n_train, n_val, length, seed=1111, ragged=False):
    """Synthetic corpus in the reference's file format: 11 meta tokens uniform in
    [560, 728], event tokens uniform in [2, 559], last event = EOS (1)."""
    os.makedirs(data_dir, exist_ok=True)
    rng = np.random.RandomState(seed)

    def make(n):
        inp = np.empty(n, dtype=object)
        tgt = np.empty(n, dtype=object)
        for i in range(n):
            ln = int(rng.randint(length // 2, 4 * length)) if ragged else length
            inp[i] = rng.randint(560, 729, size=11).astype(np.int64)
            ev = rng.randint(2, 560, size=max(2, ln - 11)).astype(np.int16)
            ev[-1] = 1
            tgt[i] = ev
        return inp, tgt

    for tag, n in (("train", n_train), ("val", n_val)):
        inp, tgt = make(n)
        np.save(os.path.join(data_dir, "input_%s.npy" % tag), inp, allow_pickle=True)
        np.save(os.path.join(data_dir, "target_%s.npy" % tag), tgt, allow_pickle=True)
