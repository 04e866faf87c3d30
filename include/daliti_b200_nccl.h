/* include/daliti_b200_nccl.h -- native NCCL transport for the sharded map (BASELINE config C4).
 *
 * The reference has no counterpart (a single-process CPU node).  This is the library-collective way of
 * summing the 158 doubles of partial normal equations (and the per-point map_incremental decisions) over the ranks that
 * hold the shards of one map: a plain ncclAllReduce(sum, double) enqueued on the handle's stream between the residual
 * pass and the solve, with no Python / torch in the process.  It plugs into the reduce-callback slot of daliti_b200.h /
 * daliti_b200_lio.h (dlt_lio_set_reduce); the peer-mailbox exchange (dlt_lio_peer_attach) is the faster alternative where
 * the GPUs have peer access.
 *
 * libnccl.so.2 is looked up at run time (dlopen): the product library has no link-time dependency on NCCL and every
 * function below returns DLT_E_STATE with a message when the library cannot be found.
 *
 *   rank 0:      dlt_nccl_unique_id(id)            -> 128 bytes, handed to every rank by the application (file, socket, MPI)
 *   every rank:  dlt_nccl_create(id, rank, world, device, &comm)
 *                dlt_nccl_attach(comm, lio)        -> dlt_lio_set_reduce(lio, dlt_nccl_allreduce, comm, NULL)
 *                ... dlt_lio_process_scan ...
 *                dlt_nccl_destroy(comm)
 */
#ifndef DALITI_B200_NCCL_H
#define DALITI_B200_NCCL_H

#include "daliti_b200_lio.h"

#ifdef __cplusplus
extern "C" {
#endif

#define DLT_NCCL_ID_BYTES 128

typedef struct dlt_nccl_comm_s *dlt_nccl_comm;

int dlt_nccl_available(void); /* 1 when libnccl.so.2 could be loaded */
const char *dlt_nccl_last_error(void);
int dlt_nccl_unique_id(unsigned char *id /* DLT_NCCL_ID_BYTES */);
int dlt_nccl_create(const unsigned char *id, int rank, int world, int device, dlt_nccl_comm *out);
int dlt_nccl_destroy(dlt_nccl_comm c);
/* a dlt_reduce_fn / dlt_lio_reduce_fn: sums n doubles at buf_dev (DEVICE memory) over the ranks in place, enqueued on the
 * stream the communicator is bound to (dlt_nccl_attach binds it to the handle's stream); never blocks the host            */
int dlt_nccl_allreduce(void *comm, double *buf_dev, int n);
/* bind the communicator to the per-scan update `lio`: its stream, and its reduce-callback slot                         */
int dlt_nccl_attach(dlt_nccl_comm c, dlt_lio lio);
/* the same for a bare device handle (dlt_set_shard_reduce; dlt_iekf_update takes the callback as an argument)           */
int dlt_nccl_attach_handle(dlt_nccl_comm c, dlt_handle h);

#ifdef __cplusplus
}
#endif
#endif /* DALITI_B200_NCCL_H */
