// qtb_nccl.cpp — the engine's own NCCL plumbing (no Python on the data path).
//
// libnccl is not a build dependency: the library torch ships (or the system's) is bound at run time with dlopen and the
// handful of entry points the engine needs are declared here (their C ABI has been stable since NCCL 2.0). The
// communicator is created from a 128-byte unique id that the host side broadcasts to the ranks once
// (quantit_b200/sharding.py does it over torch.distributed; an MPI or file-based exchange works the same).
#include <dlfcn.h>

#include <cstring>

#include "qtb_core.h"

namespace qtb
{
namespace
{
struct NcclUniqueId
{
	char internal[128];
};
using nccl_comm_t = void *;
struct NcclApi
{
	void *lib = nullptr;
	int (*GetUniqueId)(NcclUniqueId *) = nullptr;
	int (*CommInitRank)(nccl_comm_t *, int, NcclUniqueId, int) = nullptr;
	int (*CommDestroy)(nccl_comm_t) = nullptr;
	int (*AllReduce)(const void *, void *, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
	int (*Broadcast)(const void *, void *, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
	int (*AllGather)(const void *, void *, size_t, int, nccl_comm_t, cudaStream_t) = nullptr;
	int (*GroupStart)() = nullptr;
	int (*GroupEnd)() = nullptr;
	const char *(*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl; // bound once per process (dlopen handles are process-wide by nature); guarded by the caller's one-time init
constexpr int kNcclFloat64 = 8, kNcclSum = 0;

const NcclApi &nccl_api(const char *libpath)
{
	if (g_nccl.lib)
		return g_nccl;
	const char *cands[] = {libpath, "libnccl.so.2", "libnccl.so"};
	void *h = nullptr;
	for (const char *c : cands)
	{
		if (!c || !*c)
			continue;
		h = dlopen(c, RTLD_NOW | RTLD_NOLOAD); // already mapped by the host process (torch)?
		if (!h)
			h = dlopen(c, RTLD_NOW | RTLD_GLOBAL);
		if (h)
			break;
	}
	QTB_REQUIRE(h != nullptr, QTB_ERR_RUNTIME, "NCCL: libnccl.so.2 could not be loaded (pass its path to qtb_ctx_init_nccl)");
	NcclApi a;
	a.lib = h;
	auto bind = [&](const char *name)
	{
		void *p = dlsym(h, name);
		QTB_REQUIRE(p != nullptr, QTB_ERR_RUNTIME, std::string("NCCL: symbol not found: ") + name);
		return p;
	};
	a.GetUniqueId = (decltype(a.GetUniqueId))bind("ncclGetUniqueId");
	a.CommInitRank = (decltype(a.CommInitRank))bind("ncclCommInitRank");
	a.CommDestroy = (decltype(a.CommDestroy))bind("ncclCommDestroy");
	a.AllReduce = (decltype(a.AllReduce))bind("ncclAllReduce");
	a.Broadcast = (decltype(a.Broadcast))bind("ncclBroadcast");
	a.AllGather = (decltype(a.AllGather))bind("ncclAllGather");
	a.GroupStart = (decltype(a.GroupStart))bind("ncclGroupStart");
	a.GroupEnd = (decltype(a.GroupEnd))bind("ncclGroupEnd");
	a.GetErrorString = (decltype(a.GetErrorString))bind("ncclGetErrorString");
	g_nccl = a;
	return g_nccl;
}
void nccl_check(int rc, const char *what)
{
	if (rc != 0)
		throw Error(QTB_ERR_RUNTIME, std::string("NCCL: ") + what + " failed: " +
		                                 (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "unknown error"));
}
} // namespace

void nccl_unique_id(const char *libpath, char out[128])
{
	NcclUniqueId id;
	nccl_check(nccl_api(libpath).GetUniqueId(&id), "ncclGetUniqueId");
	std::memcpy(out, id.internal, 128);
}

void ctx_init_nccl(Ctx &ctx, int rank, int world, const char id_bytes[128], const char *libpath)
{
	QTB_REQUIRE(world >= 1 && rank >= 0 && rank < world, QTB_ERR_INVALID_ARGUMENT, "rank outside [0, world)");
	const NcclApi &api = nccl_api(libpath);
	if (ctx.nccl_comm)
	{
		api.CommDestroy((nccl_comm_t)ctx.nccl_comm);
		ctx.nccl_comm = nullptr;
	}
	NcclUniqueId id;
	std::memcpy(id.internal, id_bytes, 128);
	nccl_comm_t comm = nullptr;
	nccl_check(api.CommInitRank(&comm, world, id, rank), "ncclCommInitRank");
	ctx.nccl_comm = comm;
	ctx.rank = rank;
	ctx.world = world;
}

void ctx_destroy_nccl(Ctx &ctx)
{
	if (ctx.nccl_comm && g_nccl.CommDestroy)
		g_nccl.CommDestroy((nccl_comm_t)ctx.nccl_comm);
	ctx.nccl_comm = nullptr;
}

bool nccl_allreduce(Ctx &ctx, double *ptr, i64 n)
{
	if (!ctx.nccl_comm)
		return false;
	nccl_check(g_nccl.AllReduce(ptr, ptr, (size_t)n, kNcclFloat64, kNcclSum, (nccl_comm_t)ctx.nccl_comm, ctx.stream), "ncclAllReduce");
	return true;
}

// in-place all-gather of `chunk` doubles per rank: rank r's piece sits at base + r * chunk
bool nccl_allgather(Ctx &ctx, double *base, i64 chunk)
{
	if (!ctx.nccl_comm)
		return false;
	nccl_check(g_nccl.AllGather(base + (i64)ctx.rank * chunk, base, (size_t)chunk, kNcclFloat64, (nccl_comm_t)ctx.nccl_comm, ctx.stream),
	           "ncclAllGather");
	return true;
}

// every range is broadcast in place from the rank that owns it: ONE grouped NCCL operation, exactly the owned bytes move
bool nccl_exchange_ranges(Ctx &ctx, double *base, const std::vector<OwnedRange> &ranges)
{
	if (!ctx.nccl_comm)
		return false;
	nccl_check(g_nccl.GroupStart(), "ncclGroupStart");
	for (const OwnedRange &r : ranges)
		if (r.n > 0)
			nccl_check(g_nccl.Broadcast(base + r.off, base + r.off, (size_t)r.n, kNcclFloat64, r.owner, (nccl_comm_t)ctx.nccl_comm, ctx.stream),
			           "ncclBroadcast");
	nccl_check(g_nccl.GroupEnd(), "ncclGroupEnd");
	return true;
}

} // namespace qtb
