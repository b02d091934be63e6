"""Host-resident execution of a ConvolutionPlan: upload, convolve and read back in row chunks on three streams.

When features live in (pinned) host memory, PCIe is the bound: one forward + backward of the 64-channel bench
workload moves ~420 MB each way but computes for only ~1.4 ms.  ``HostPipelinedConv`` cuts the batch into
tile-aligned row chunks and overlaps the three phases (PCIe is full duplex): while chunk c is convolved, chunk
c+1 is uploading and chunk c-1 is being read back.  Every chunk goes through the same C-ABI kernels
(``fvc_conv_forward`` / ``fvc_conv_wgrad``) on a sub-range of output rows -- the dense tap-major map, its tile
masks and the outputs are simply offset by the chunk's first row (a multiple of 128).

Requires a same-topology plan (source and target grid identical, so feature rows and output rows coincide) and a
dtype / channel combination served by the tensor-core kernels (the chunked weight gradient uses the dense map).
"""

from __future__ import annotations

import ctypes as C

import torch

from . import _fvdb_cpp as cpp
from ._lib import check, lib
from .convolution_plan import ConvolutionPlan, _GatherScatterBackend


class HostPipelinedConv:
    def __init__(self, plan: ConvolutionPlan, num_chunks: int = 8):
        if not isinstance(plan._backend, _GatherScatterBackend):
            raise ValueError("HostPipelinedConv needs a kernel-map (gather-scatter) plan")
        if not plan.has_fixed_topology:
            raise ValueError("HostPipelinedConv needs a same-topology plan (target_grid is source_grid)")
        self.plan, self.topo = plan, plan._backend.topology
        self.device = self.topo.device
        n = self.topo.output_total_voxels
        tiles = (n + 127) // 128
        per = max(1, (tiles + num_chunks - 1) // num_chunks)
        self.bounds = [(t * 128, min((t + per) * 128, n)) for t in range(0, tiles, per)]
        self.s_in, self.s_out = torch.cuda.Stream(self.device), torch.cuda.Stream(self.device)
        self._ws = {}
        # upload dependencies, read off the map itself: chunk c may start once the chunk holding the largest row it
        # gathers has landed (neighbours never leave the grid, but a grid can span any number of chunks)
        self.fwd_needs = self._last_chunk_needed(self.topo._out_map())
        self.bwd_needs = self._last_chunk_needed(self.topo._dgrad_plan()[0])

    def _last_chunk_needed(self, nbr: torch.Tensor) -> list[int]:
        tops = torch.stack([nbr[:, r0:r1].amax() for r0, r1 in self.bounds]).tolist()  # one sync, at construction
        starts = [r0 for r0, _ in self.bounds]
        needs = []
        for c, top in enumerate(tops):
            holder = max(i for i, r0 in enumerate(starts) if r0 <= max(int(top), 0))
            needs.append(max(c, holder))
        return needs

    # ---- persistent workspace: device buffers, scratch, events and per-chunk call arguments --------------------------
    def _workspace(self, dtype, cin, cout, weight_shape):
        key = (dtype, cin, cout)
        ws = self._ws.get(key)
        if ws is not None:
            return ws
        topo, dev = self.topo, self.device
        n, k3, code = topo.output_total_voxels, topo.kernel_volume, cpp._DTYPE_CODE[dtype]
        chunks = len(self.bounds)
        rows = max(r1 - r0 for r0, r1 in self.bounds)
        ws = {
            "x": torch.empty((n, cin), dtype=dtype, device=dev), "dy": torch.empty((n, cout), dtype=dtype, device=dev),
            "y": torch.empty((n, cout), dtype=dtype, device=dev), "gx": torch.empty((n, cin), dtype=dtype, device=dev),
            "gw_chunks": torch.empty((chunks, *weight_shape), dtype=dtype, device=dev),
            "fwd_bytes": int(lib.fvc_conv_scratch_bytes(n, rows, cin, cout, k3, code)),
            "bwd_bytes": int(lib.fvc_conv_scratch_bytes(n, rows, cout, cin, k3, code)),
            "wg_bytes": int(lib.fvc_conv_wgrad_scratch_bytes(n, rows, 0, cin, cout, k3, code, 2, 1)),
            "x_ready": [torch.cuda.Event() for _ in range(chunks)], "dy_ready": [torch.cuda.Event() for _ in range(chunks)],
            "y_done": [torch.cuda.Event() for _ in range(chunks)], "g_done": [torch.cuda.Event() for _ in range(chunks)],
            "final": torch.cuda.Event(),
        }
        # one scratch per kernel family: the chunk calls of a family run back to back on the calling stream
        ws["fwd_scratch"] = torch.empty(max(ws["fwd_bytes"], 16), dtype=torch.uint8, device=dev)
        ws["bwd_scratch"] = torch.empty(max(ws["bwd_bytes"], 16), dtype=torch.uint8, device=dev)
        ws["wg_scratch"] = torch.empty(max(ws["wg_bytes"], 16), dtype=torch.uint8, device=dev)
        self._ws[key] = ws
        return ws

    def _conv_rows(self, x, w_prepared, nbr, mask, r0, r1, cin, cout, out, scratch, scratch_bytes, stream):
        """Rows [r0, r1) of the output-stationary kernel: map, tile mask and output are offset by the chunk's first row."""
        k3, pitch = int(nbr.shape[0]), int(nbr.stride(0))
        words = (k3 + 63) // 64
        check(
            lib.fvc_conv_forward_ex(
                x.data_ptr(), 0, w_prepared.data_ptr(), None, out.data_ptr() + r0 * cout * out.element_size(), nbr.data_ptr() + 4 * r0, pitch,
                (mask.data_ptr() + 8 * words * (r0 // 128)) if mask is not None else None, int(x.shape[0]), r1 - r0, cin, cout, k3,
                cpp._DTYPE_CODE[x.dtype], 2, scratch.data_ptr(), scratch_bytes, stream,
            )
        )

    def _wgrad_rows(self, x, dy, r0, r1, cin, cout, grad_w, scratch, scratch_bytes, stream):
        topo = self.topo
        nbr, mask = topo._out_map(), topo._out_mask()
        k3, pitch = int(nbr.shape[0]), int(nbr.stride(0))
        words = (k3 + 63) // 64
        check(
            lib.fvc_conv_wgrad(
                x.data_ptr(), dy.data_ptr() + r0 * cout * dy.element_size(), None, None, None, None, nbr.data_ptr() + 4 * r0, pitch,
                (mask.data_ptr() + 8 * words * (r0 // 128)) if mask is not None else None, int(x.shape[0]), r1 - r0, cin, cout, k3,
                cpp._DTYPE_CODE[x.dtype], 2, grad_w.data_ptr(), scratch.data_ptr(), scratch_bytes, stream,
            )
        )

    # ---- the pipelined step ----------------------------------------------------------------------
    def forward_backward(self, x_host, dy_host, weights, y_host, gx_host, gw_host, reduce_fn=None):
        """y = conv(x), (gx, gw) = conv_backward(dy) with x / dy read from and y / gx / gw written to pinned host tensors.
        ``reduce_fn(gw)`` (e.g. an NCCL all-reduce) runs on the weight gradient before it is read back.  The calling
        stream waits for the read-back stream, so the step is complete when this stream is.  Device buffers, scratch and
        events persist between calls (stream order protects their reuse), so a step costs two C-ABI calls per chunk and
        phase plus the copies -- the host never becomes the bound."""
        topo, dev = self.topo, self.device
        dtype = weights.dtype
        cout, cin = int(weights.shape[0]), int(weights.shape[1])
        ws = self._workspace(dtype, cin, cout, tuple(weights.shape))
        x, dy, y, gx, gw_chunks = ws["x"], ws["dy"], ws["y"], ws["gx"], ws["gw_chunks"]
        main = torch.cuda.current_stream(dev)
        stream = main.cuda_stream
        in_map, in_mask, mirror = topo._dgrad_plan()
        out_map, out_mask = topo._out_map(), topo._out_mask()
        with torch.cuda.device(dev):
            prev_path, cpp._path = cpp._path, 2  # the chunk calls force the tensor-core family: prepare its operand
            try:  # ONE weight image per direction and step, shared by every chunk
                w_fwd, w_bwd = cpp._prepare_weights(weights, dtype, False), cpp._prepare_weights(weights, dtype, True, flip_taps=mirror)
            finally:
                cpp._path = prev_path
            self.s_in.wait_stream(main)  # the previous step's kernels are done with x / dy before they are overwritten
            self.s_out.wait_stream(main)
            with torch.cuda.stream(self.s_in):  # uploads in the order the kernels need them
                for c, (r0, r1) in enumerate(self.bounds):
                    x[r0:r1].copy_(x_host[r0:r1], non_blocking=True)
                    ws["x_ready"][c].record(self.s_in)
                for c, (r0, r1) in enumerate(self.bounds):
                    dy[r0:r1].copy_(dy_host[r0:r1], non_blocking=True)
                    ws["dy_ready"][c].record(self.s_in)
            for c, (r0, r1) in enumerate(self.bounds):  # forward
                main.wait_event(ws["x_ready"][self.fwd_needs[c]])
                self._conv_rows(x, w_fwd, out_map, out_mask, r0, r1, cin, cout, y, ws["fwd_scratch"], ws["fwd_bytes"], stream)
                ws["y_done"][c].record(main)
            with torch.cuda.stream(self.s_out):
                for c, (r0, r1) in enumerate(self.bounds):
                    self.s_out.wait_event(ws["y_done"][c])
                    y_host[r0:r1].copy_(y[r0:r1], non_blocking=True)
            for c, (r0, r1) in enumerate(self.bounds):  # backward
                main.wait_event(ws["dy_ready"][self.bwd_needs[c]])
                self._wgrad_rows(x, dy, r0, r1, cin, cout, gw_chunks[c], ws["wg_scratch"], ws["wg_bytes"], stream)
                self._conv_rows(dy, w_bwd, in_map, in_mask, r0, r1, cout, cin, gx, ws["bwd_scratch"], ws["bwd_bytes"], stream)
                ws["g_done"][c].record(main)
            with torch.cuda.stream(self.s_out):
                for c, (r0, r1) in enumerate(self.bounds):
                    self.s_out.wait_event(ws["g_done"][c])
                    gx_host[r0:r1].copy_(gx[r0:r1], non_blocking=True)
            gw = gw_chunks.float().sum(dim=0).to(dtype)  # chunk partials summed in fp32, rounded once
            if reduce_fn is not None:
                reduce_fn(gw)
            ws["final"].record(main)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(ws["final"])
                gw_host.copy_(gw, non_blocking=True)
            gw.record_stream(self.s_out)
            main.wait_stream(self.s_out)
        return gw
