"""CPU emulation of the multi-GPU slab protocol (test infrastructure).

Every rank steps the WHOLE grid with the CPU restatement but only trusts the rows it owns plus
`halo` rows either side: everything else is overwritten with a huge finite poison value after every
exchange, so if the halo width or the set of exchanged rows were insufficient, the owned rows would
differ from the single-domain run. (The poison is finite, not NaN, because the reference multiplies
rows beyond the lossless stencil by stored ZERO coefficients of a_vx_vx: 0 * poison must stay 0.) After every
`k` steps the outermost `halo` owned rows of all components travel to the neighbours through
torch.distributed (gloo on CPU) -- the same rows `libfdsb200.so` sends with ncclSend/ncclRecv.
"""

import numpy as np
import torch
import torch.distributed as dist

from oracle import restate
from pyfds_b200 import parallel


POISON = 1e30


def run_slab(field, steps, rank, world, steps_per_exchange):
    nx, ny = field.x.samples, field.y.samples
    row0, rows = parallel.partition_rows(ny, world)[rank]
    halo = steps_per_exchange * parallel.stencil_reach(field)
    stepper = restate.stepper_for(field)
    names = stepper.components
    lo, hi = max(row0 - halo, 0), min(row0 + rows + halo, ny)

    def poison():
        for name in names:
            v = stepper.values(name).reshape(ny, nx)
            v[:lo] = POISON
            v[hi:] = POISON

    def exchange():
        for name in names:
            v = stepper.values(name).reshape(ny, nx)
            ops = []
            if rank > 0:
                send = torch.from_numpy(v[row0:row0 + halo].copy())
                recv = torch.empty((row0 - lo, nx), dtype=torch.float64)
                ops += [dist.P2POp(dist.isend, send, rank - 1), dist.P2POp(dist.irecv, recv, rank - 1)]
            if rank < world - 1:
                send2 = torch.from_numpy(v[row0 + rows - halo:row0 + rows].copy())
                recv2 = torch.empty((hi - row0 - rows, nx), dtype=torch.float64)
                ops += [dist.P2POp(dist.isend, send2, rank + 1),
                        dist.P2POp(dist.irecv, recv2, rank + 1)]
            for req in dist.batch_isend_irecv(ops):
                req.wait()
            if rank > 0:
                v[lo:row0] = recv.numpy()[halo - (row0 - lo):]
            if rank < world - 1:
                v[row0 + rows:hi] = recv2.numpy()[:hi - row0 - rows]

    poison()
    done = 0
    while done < steps:
        count = min(steps_per_exchange, steps - done)
        with np.errstate(all='ignore'):
            stepper.run(count)
        done += count
        exchange()
        poison()
    owned = {name: stepper.values(name).reshape(ny, nx)[row0:row0 + rows].copy() for name in names}
    return row0, rows, owned, stepper
