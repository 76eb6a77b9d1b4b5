"""Minimal single-read fast5 (HDF5) reader for the raw basecall path -- no h5py required.

The reference reads raw signal through the third-party `fast5_research.Fast5` wrapper
(`sloika/basecall.py:103-106`: `Fast5(fn).get_read(raw=True)`, `.filename_short`;
`bin/basecall_network.py:96`: `iterate_fast5(folder, paths=True, limit, strand_list)`), which sits
on h5py.  Neither is available here, so this module parses the subset of HDF5 that MinKNOW-era
single-read fast5 files use:

  superblock v0/v1 . v1 object headers (+ continuation blocks) . symbol-table groups (v1 B-tree +
  local heap + SNOD) . contiguous / compact / chunked (v1 chunk B-tree) dataset layouts .
  filter pipeline v1/v2 with shuffle (2) and deflate (1) . fixed-point / IEEE float / fixed-length
  string datatypes for datasets and attributes (attribute messages v1-v3).

`Fast5.get_read(raw=True)` returns the signal scaled to pA exactly as fast5_research does:
`(daq + offset) * range / digitisation` with the `UniqueGlobalKey/channel_id` attributes.
Raw lengths of the bundled reads are pinned by the reference's `test/unit/test_fast5.py:99-110`.
"""
import os
import zlib

import numpy as np

_SIG = b'\x89HDF\r\n\x1a\n'
_UNDEF = 0xFFFFFFFFFFFFFFFF


class Fast5Error(IOError):
    pass


class _Dataset(object):
    def __init__(self):
        self.shape = None
        self.dtype = None
        self.layout = None      # ('contiguous', addr, size) | ('compact', bytes) | ('chunked', btree, chunk_dims)
        self.filters = []       # [(id, client_data)]
        self.attrs = {}


class _Obj(object):
    """Parsed object header: either a group (has `symtab`) or a dataset."""

    def __init__(self):
        self.symtab = None      # (btree_addr, heap_addr)
        self.data = _Dataset()
        self.attrs = {}


class H5File(object):
    def __init__(self, filename):
        self.filename = filename
        with open(filename, 'rb') as fh:
            self.buf = fh.read()
        self._parse_superblock()
        self._cache = {}

    # -- primitives -----------------------------------------------------------------------------
    def _u(self, off, size):
        return int.from_bytes(self.buf[off:off + size], 'little')

    def _parse_superblock(self):
        b = self.buf
        base = b.find(_SIG)
        if base != 0:
            raise Fast5Error("{}: not an HDF5 file".format(self.filename))
        version = b[8]
        if version not in (0, 1):
            raise Fast5Error("{}: HDF5 superblock version {} not supported".format(self.filename, version))
        self.O = b[13]
        self.L = b[14]
        pos = 24 if version == 0 else 28
        pos += 4 * self.O       # base, free-space, eof, driver addresses
        # root symbol table entry
        self.root_header = self._u(pos + self.O, self.O)

    # -- object headers -------------------------------------------------------------------------
    def _messages(self, addr):
        b = self.buf
        version = b[addr]
        if version != 1:
            raise Fast5Error("object header version {} not supported".format(version))
        nmsg = self._u(addr + 2, 2)
        size = self._u(addr + 8, 4)
        blocks = [(addr + 16, size)]
        out = []
        while blocks and len(out) < nmsg:
            pos, remaining = blocks.pop(0)
            end = pos + remaining
            while pos + 8 <= end and len(out) < nmsg:
                mtype = self._u(pos, 2)
                msize = self._u(pos + 2, 2)
                body = pos + 8
                if mtype == 0x0010:
                    blocks.append((self._u(body, self.O), self._u(body + self.O, self.L)))
                out.append((mtype, body, msize))
                pos = body + msize
        return out

    def _object(self, addr):
        if addr in self._cache:
            return self._cache[addr]
        obj = _Obj()
        for mtype, body, msize in self._messages(addr):
            if mtype == 0x0011:
                obj.symtab = (self._u(body, self.O), self._u(body + self.O, self.O))
            elif mtype == 0x0001:
                obj.data.shape = self._dataspace(body)[0]
            elif mtype == 0x0003:
                try:
                    obj.data.dtype = self._datatype(body)[0]
                except Fast5Error:
                    obj.data.dtype = None   # compound/vlen tables (event data) are not on this path
            elif mtype == 0x0008:
                obj.data.layout = self._layout(body)
            elif mtype == 0x000B:
                obj.data.filters = self._filters(body)
            elif mtype == 0x000C:
                try:
                    name, value = self._attribute(body)
                    obj.attrs[name] = value
                except Fast5Error:
                    pass
        obj.data.attrs = obj.attrs
        self._cache[addr] = obj
        return obj

    def _dataspace(self, pos):
        b = self.buf
        version, rank, flags = b[pos], b[pos + 1], b[pos + 2]
        if version == 1:
            p = pos + 8
        elif version == 2:
            p = pos + 4
        else:
            raise Fast5Error("dataspace version {}".format(version))
        dims = tuple(self._u(p + i * self.L, self.L) for i in range(rank))
        p += rank * self.L
        if flags & 1:
            p += rank * self.L
        return dims, p - pos

    def _datatype(self, pos):
        b = self.buf
        cls = b[pos] & 0x0F
        bits0 = b[pos + 1]
        size = self._u(pos + 4, 4)
        if cls == 0:        # fixed point
            order = '>' if bits0 & 1 else '<'
            kind = 'i' if bits0 & 8 else 'u'
            return np.dtype('{}{}{}'.format(order, kind, size)), 8 + 4
        if cls == 1:        # floating point
            order = '>' if bits0 & 1 else '<'
            return np.dtype('{}f{}'.format(order, size)), 8 + 12
        if cls == 3:        # fixed-length string
            return np.dtype('S{}'.format(size)), 8
        if cls == 8:        # enumeration (h5py stores bool arrays as an int8 enum): the base type is what is stored
            nmemb = bits0 | (b[pos + 2] << 8)
            base, used = self._datatype(pos + 8)
            cur = pos + 8 + used
            version = b[pos] >> 4
            for _ in range(nmemb):
                end = b.index(b'\x00', cur)
                cur = cur + ((end - cur) // 8 + 1) * 8 if version < 3 else end + 1
            return base, cur + nmemb * base.itemsize - pos
        if cls == 6:        # compound (event tables): versions 1-3 of the datatype message
            version = b[pos] >> 4
            nmemb = bits0 | (b[pos + 2] << 8)
            cur = pos + 8
            names, formats, offsets = [], [], []
            for _ in range(nmemb):
                end = b.index(b'\x00', cur)
                names.append(bytes(b[cur:end]).decode('ascii'))
                if version < 3:
                    cur += ((end - cur) // 8 + 1) * 8                    # name null padded to a multiple of 8
                    offsets.append(self._u(cur, 4))
                    cur += 4
                    if version == 1:
                        cur += 1 + 3 + 4 + 4 + 16                         # dimensionality, reserved, permutation, reserved, sizes
                else:
                    cur = end + 1
                    nb = max(1, (int(size).bit_length() + 7) // 8)
                    offsets.append(self._u(cur, nb))
                    cur += nb
                sub, used = self._datatype(cur)
                formats.append(sub)
                cur += used
            return np.dtype({'names': names, 'formats': formats, 'offsets': offsets, 'itemsize': size}), cur - pos
        raise Fast5Error("datatype class {} not supported".format(cls))

    def _layout(self, pos):
        b = self.buf
        version = b[pos]
        if version != 3:
            raise Fast5Error("data layout version {} not supported".format(version))
        cls = b[pos + 1]
        if cls == 0:
            size = self._u(pos + 2, 2)
            return ('compact', bytes(b[pos + 4:pos + 4 + size]))
        if cls == 1:
            return ('contiguous', self._u(pos + 2, self.O), self._u(pos + 2 + self.O, self.L))
        if cls == 2:
            rank = b[pos + 2]
            btree = self._u(pos + 3, self.O)
            dims = tuple(self._u(pos + 3 + self.O + 4 * i, 4) for i in range(rank))
            return ('chunked', btree, dims)
        raise Fast5Error("layout class {}".format(cls))

    def _filters(self, pos):
        b = self.buf
        version, nfilt = b[pos], b[pos + 1]
        out = []
        p = pos + (8 if version == 1 else 2)
        for _ in range(nfilt):
            fid = self._u(p, 2)
            if version == 1 or fid >= 256:
                namelen = self._u(p + 2, 2)
                p += 4
            else:
                namelen = 0
                p += 2
            ncd = self._u(p + 2, 2)
            p += 4
            if version == 1:
                namelen = (namelen + 7) // 8 * 8
            p += namelen
            cd = [self._u(p + 4 * i, 4) for i in range(ncd)]
            p += 4 * ncd
            if version == 1 and ncd % 2:
                p += 4
            out.append((fid, cd))
        return out

    def _attribute(self, pos):
        b = self.buf
        version = b[pos]
        nsize, tsize, ssize = self._u(pos + 2, 2), self._u(pos + 4, 2), self._u(pos + 6, 2)
        if version == 1:
            pad = lambda n: (n + 7) // 8 * 8
            p = pos + 8
        elif version in (2, 3):
            pad = lambda n: n
            p = pos + (8 if version == 2 else 9)
        else:
            raise Fast5Error("attribute version {}".format(version))
        name = bytes(b[p:p + nsize]).split(b'\0')[0].decode('utf-8', 'replace')
        p += pad(nsize)
        dtype, _ = self._datatype(p)
        p += pad(tsize)
        shape, _ = self._dataspace(p) if ssize >= 4 else ((), 0)
        p += pad(ssize)
        count = int(np.prod(shape)) if len(shape) else 1
        raw = bytes(b[p:p + count * dtype.itemsize])
        arr = np.frombuffer(raw, dtype=dtype, count=count)
        if dtype.kind == 'S':
            vals = [v.split(b'\0')[0] for v in arr]
            value = vals[0] if not shape else np.array(vals)
        else:
            value = arr[0] if not shape else arr.reshape(shape).copy()
        return name, value

    # -- groups ---------------------------------------------------------------------------------
    def _heap_string(self, heap_addr, offset):
        if self.buf[heap_addr:heap_addr + 4] != b'HEAP':
            raise Fast5Error("bad local heap")
        data = self._u(heap_addr + 8 + 2 * self.L, self.O)
        start = data + offset
        end = self.buf.index(b'\0', start)
        return bytes(self.buf[start:end]).decode('utf-8', 'replace')

    def _group_entries(self, btree, heap):
        out = {}
        sig = self.buf[btree:btree + 4]
        if sig == b'SNOD':
            nsym = self._u(btree + 6, 2)
            p = btree + 8
            for _ in range(nsym):
                name = self._heap_string(heap, self._u(p, self.O))
                out[name] = self._u(p + self.O, self.O)
                p += 2 * self.O + 24
            return out
        if sig != b'TREE':
            raise Fast5Error("bad group B-tree node")
        used = self._u(btree + 6, 2)
        p = btree + 8 + 2 * self.O
        for i in range(used):
            child = self._u(p + self.L, self.O)
            out.update(self._group_entries(child, heap))
            p += self.L + self.O
        return out

    def children(self, path='/'):
        obj = self._resolve(path)
        if obj.symtab is None:
            raise Fast5Error("{} is not a group".format(path))
        return self._group_entries(*obj.symtab)

    def _resolve(self, path):
        addr = self.root_header
        for part in [p for p in path.split('/') if p]:
            obj = self._object(addr)
            if obj.symtab is None:
                raise KeyError(path)
            entries = self._group_entries(*obj.symtab)
            if part not in entries:
                raise KeyError(path)
            addr = entries[part]
        return self._object(addr)

    def __contains__(self, path):
        try:
            self._resolve(path)
            return True
        except KeyError:
            return False

    def attrs(self, path):
        return self._resolve(path).attrs

    # -- datasets -------------------------------------------------------------------------------
    def _chunks(self, node, rank):
        """Yield (offsets, size, filter_mask, address) from a v1 chunk B-tree."""
        if node == _UNDEF:
            return
        if self.buf[node:node + 4] != b'TREE':
            raise Fast5Error("bad chunk B-tree node")
        level = self.buf[node + 5]
        used = self._u(node + 6, 2)
        keysize = 8 + 8 * rank
        p = node + 8 + 2 * self.O
        for _ in range(used):
            size = self._u(p, 4)
            mask = self._u(p + 4, 4)
            offs = tuple(self._u(p + 8 + 8 * i, 8) for i in range(rank - 1))
            child = self._u(p + keysize, self.O)
            if level == 0:
                yield offs, size, mask, child
            else:
                for item in self._chunks(child, rank):
                    yield item
            p += keysize + self.O

    @staticmethod
    def _unfilter(raw, filters, mask, itemsize):
        for idx in range(len(filters) - 1, -1, -1):
            if mask & (1 << idx):
                continue
            fid, _ = filters[idx]
            if fid == 1:
                raw = zlib.decompress(raw)
            elif fid == 2:
                n = len(raw) // itemsize
                arr = np.frombuffer(raw, dtype=np.uint8, count=n * itemsize)
                if itemsize == 2:
                    # the raw DAQ signal (int16): byte planes recombined with contiguous vector operations, an
                    # order of magnitude faster than the strided transpose
                    words = arr[:n].astype('<u2')
                    words |= arr[n:].astype('<u2') << 8
                    raw = words.tobytes() + raw[n * itemsize:]
                else:
                    raw = arr.reshape(itemsize, n).T.tobytes() + raw[n * itemsize:]
            elif fid == 3:
                raw = raw[:-4]          # fletcher32 checksum trailer
            else:
                raise Fast5Error("HDF5 filter {} not supported".format(fid))
        return raw

    def dataset(self, path):
        ds = self._resolve(path).data
        if ds.shape is None or ds.dtype is None or ds.layout is None:
            raise Fast5Error("{} is not a dataset".format(path))
        kind = ds.layout[0]
        count = int(np.prod(ds.shape)) if len(ds.shape) else 1
        if kind == 'compact':
            return np.frombuffer(ds.layout[1], dtype=ds.dtype, count=count).reshape(ds.shape).copy()
        if kind == 'contiguous':
            addr = ds.layout[1]
            if addr == _UNDEF:
                return np.zeros(ds.shape, dtype=ds.dtype)
            raw = self.buf[addr:addr + count * ds.dtype.itemsize]
            return np.frombuffer(raw, dtype=ds.dtype, count=count).reshape(ds.shape).copy()
        _, btree, cdims = ds.layout
        rank = len(cdims)
        if rank - 1 != len(ds.shape):
            raise Fast5Error("chunk rank mismatch")
        out = np.zeros(ds.shape, dtype=ds.dtype)
        cshape = cdims[:-1]
        for offs, size, mask, addr in self._chunks(btree, rank):
            raw = self._unfilter(bytes(self.buf[addr:addr + size]), ds.filters, mask, ds.dtype.itemsize)
            chunk = np.frombuffer(raw, dtype=ds.dtype, count=int(np.prod(cshape))).reshape(cshape)
            sel_out = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cshape, ds.shape))
            sel_in = tuple(slice(0, s.stop - s.start) for s in sel_out)
            out[sel_out] = chunk[sel_in]
        return out


class Fast5(object):
    """The part of `fast5_research.Fast5` that `sloika/basecall.py:103-106` uses."""

    def __init__(self, filename):
        self.filename = filename
        self._h5 = H5File(filename)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    @property
    def filename_short(self):
        return os.path.splitext(os.path.basename(self.filename))[0]

    def _read_group(self):
        reads = self._h5.children('/Raw/Reads')
        if not reads:
            raise Fast5Error("{}: no raw reads".format(self.filename))
        name = sorted(reads)[0]
        return '/Raw/Reads/' + name

    @property
    def channel_meta(self):
        return self._h5.attrs('/UniqueGlobalKey/channel_id')

    def get_read(self, raw=False, scale=True):
        """Raw signal of the (single) read.  `raw=True` is the only mode on this path.

        :param scale: convert DAQ integers to pA, `(daq + offset) * range / digitisation`
        """
        if not raw:
            raise NotImplementedError("only raw=True (raw signal) is supported on the B200 path")
        signal = self._h5.dataset(self._read_group() + '/Signal')
        if not scale:
            return signal
        meta = self.channel_meta
        raw_unit = float(meta['range']) / float(meta['digitisation'])
        return (signal + float(meta['offset'])) * raw_unit


    def _event_detection_group(self):
        groups = sorted(g for g in self._h5.children('/Analyses') if g.startswith('EventDetection'))
        if not groups:
            raise Fast5Error("{}: no EventDetection analysis".format(self.filename))
        base = '/Analyses/' + groups[-1] + '/Reads'
        reads = sorted(self._h5.children(base))
        if not reads:
            raise Fast5Error("{}: no event-detection reads".format(self.filename))
        return base + '/' + reads[0]

    def get_section_events(self, section, analysis='Segmentation'):
        """Events of the 'template' or 'complement' section of the read, as `fast5_research.Fast5.get_section_events`
        gives them to `basecall.events_worker` (`sloika/basecall.py:75-77`): the event-detection table
        (`/Analyses/EventDetection_*/Reads/Read_*/Events`, fields start / length / mean / stdv) cut to the section
        named by the segmentation summary (`/Analyses/<analysis>_*/Summary/<...>`), which gives either event indices
        (`start_index_temp`, `end_index_temp`, ...) or sample positions (`first_sample_template`,
        `duration_template`, ...).  fast5_research (>= 1.2.2, `requirements.txt:5`) is not vendored with the
        reference and the bundled reads carry no event-detection group: restated from that package's documented
        behaviour, exercised with synthetic files only."""
        if section not in ('template', 'complement'):
            raise ValueError("section must be 'template' or 'complement'")
        grp = self._event_detection_group()
        events = self._h5.dataset(grp + '/Events')
        segs = sorted(g for g in self._h5.children('/Analyses') if g.startswith(analysis))
        if not segs:
            raise Fast5Error("{}: no {} analysis".format(self.filename, analysis))
        summary = '/Analyses/' + segs[-1] + '/Summary'
        attrs = {}
        for child in self._h5.children(summary):
            attrs.update(self._h5.attrs(summary + '/' + child))
        short = 'temp' if section == 'template' else 'comp'
        if 'start_index_' + short in attrs:
            lo, hi = int(attrs['start_index_' + short]), int(attrs['end_index_' + short])
            return events[lo:hi + 1]
        if 'first_sample_' + section in attrs:
            first, dur = int(attrs['first_sample_' + section]), int(attrs['duration_' + section])
            start = events['start'].astype(np.int64)
            start = start - (int(self._h5.attrs(grp).get('start_time', 0)) if start.size and start[0] > first + dur else 0)
            keep = (start >= first) & (start < first + dur)
            return events[keep]
        raise Fast5Error("{}: segmentation summary names no {} section".format(self.filename, section))


def iterate_fast5(path, strand_list=None, paths=False, mode='r', limit=None, files_group_pattern=None,
                  sort_by_size=None):
    """Yield fast5 file names under `path` (non-recursive `*.fast5`), optionally restricted to the
    `filename` column of a tab-separated strand list -- the subset of
    `fast5_research.iterate_fast5` that `bin/basecall_network.py:96` relies on (`paths=True`)."""
    if not paths:
        raise NotImplementedError("only paths=True is supported")
    if strand_list is None:
        names = sorted(f for f in os.listdir(path) if f.endswith('.fast5'))
    else:
        with open(strand_list, 'r') as fh:
            header = fh.readline().rstrip('\n').split('\t')
            col = header.index('filename') if 'filename' in header else 0
            names = [line.rstrip('\n').split('\t')[col] for line in fh if line.strip()]
    if limit is not None:
        names = names[:limit]
    for name in names:
        yield os.path.join(path, name)
