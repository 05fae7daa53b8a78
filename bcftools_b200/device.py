"""Device-resident batches: torch tensors as the owner of HBM, raw pointers handed to the C-ABI.

PyTorch is plumbing here (device memory, streams, torch.distributed); all arithmetic is in the CUDA library.
"""
import numpy as np
import torch

from . import abi


def _t(a, device):
    if a is None:
        return None
    a = np.ascontiguousarray(a)
    if a.dtype == np.uint16:      # torch has limited uint16 support: ship as int16 bits
        a = a.view(np.int16)
    if a.dtype == np.uint32:
        a = a.view(np.int32)
    return torch.from_numpy(a).to(device, non_blocking=False)


class DeviceBatch:
    def __init__(self, host_batch, device="cuda:0", replicate=1):
        """Upload a HostBatch.  replicate>1 tiles the same sites `replicate` times in HBM (distinct addresses,
        same bytes) to build a resident job larger than what the numpy generator produces in reasonable time."""
        hb = host_batch
        self.host = hb
        self.device = torch.device(device)
        self.nsmpl, self.max_nals = hb.nsmpl, hb.max_nals
        self.replicate = int(replicate)
        r = self.replicate
        self.nsites = hb.nsites * r
        ext = int(hb.pl.size)
        self.t = {}
        pl = _t(hb.pl, self.device)
        self.t["pl"] = pl if r == 1 else pl.repeat(r)
        off = hb.pl_off if r == 1 else np.concatenate([hb.pl_off + k * ext for k in range(r)])
        self.pl_off_host = off
        self.t["pl_off"] = _t(off, self.device)
        for name in ("nals", "unseen", "ploidy_id", "nqs", "prior_an"):
            v = getattr(hb, name)
            self.t[name] = None if v is None else _t(np.tile(v, r), self.device)
        for name in ("qs", "prior_ac"):
            v = getattr(hb, name)
            self.t[name] = None if v is None else _t(np.tile(v, (r, 1)), self.device)
        self.t["ad"] = self.t["ad_off"] = self.t["nad"] = None
        if hb.ad is not None:
            aext = int(hb.ad.size)
            ad = _t(hb.ad, self.device)
            self.t["ad"] = ad if r == 1 else ad.repeat(r)
            self.t["ad_off"] = _t(hb.ad_off if r == 1 else np.concatenate([hb.ad_off + k * aext for k in range(r)]), self.device)
            self.t["nad"] = _t(np.tile(hb.nad, r), self.device)

    def c_struct(self):
        b = abi.McbBatch()
        b.nsites = self.nsites
        for name in abi.BATCH_FIELDS:
            v = self.t.get(name)
            setattr(b, name, None if v is None else v.data_ptr())
        b.pl_type = getattr(self.host, "pl_type", 0)
        return b

    def pl_bytes(self):
        return self.t["pl"].numel() * 4


class DeviceResult:
    def __init__(self, dbatch, want_gq=True, want_gp=False):
        d, R, S, M = dbatch.device, dbatch.nsites, dbatch.nsmpl, dbatch.max_nals
        self.dbatch = dbatch
        z = lambda shape, dt: torch.zeros(shape, dtype=dt, device=d)
        self.t = dict(
            ret=z(R, torch.int32), als_new=z(R, torch.int32), als_map=z((R, M), torch.int8), qual=z(R, torch.float32),
            ac=z((R, M), torch.int32), an=z(R, torch.int32), site_flags=z(R, torch.int32), diag=z((R, 4), torch.float64),
            gt=z((R, S, 2), torch.int32), gq=z((R, S), torch.int32) if want_gq else None,
            gp=z(dbatch.t["pl"].numel(), torch.float32) if want_gp else None,
            pl=z(dbatch.t["pl"].numel(), torch.int32))

    def c_struct(self):
        r = abi.McbResult()
        for name in abi.RESULT_FIELDS:
            v = self.t.get(name)
            setattr(r, name, None if v is None else v.data_ptr())
        return r

    def to_host(self, first_sites=None):
        """Download into an abi.HostResult (only valid for replicate==1 or the first copy)."""
        hb = self.dbatch.host
        n = hb.nsites if first_sites is None else first_sites
        res = abi.HostResult(hb, want_gp=self.t["gp"] is not None)
        for name, dt in abi.RESULT_FIELDS.items():
            v = self.t.get(name)
            if v is None:
                continue
            a = v.cpu().numpy()
            if name in ("pl", "gp"):
                a = a[:hb.pl.size]
            else:
                a = a[:hb.nsites]
            getattr(res, name)[...] = a.view(dt).reshape(getattr(res, name).shape)
        return res
