# RaycoreCUDA.jl — Julia shim binding libraycore_cuda.so (include/raycore_cuda.h) behind Raycore's own accel API.
#
# NOT EXECUTED IN THIS REPO'S CI: neither the build container nor the B200 boxes have Julia.  The shim is kept
# deliberately thin (every method is one ccall plus bookkeeping that mirrors src/instanced-bvh.jl) so it can be
# checked by inspection; every behaviour it relies on is exercised through the same C ABI by the Python mirror
# (raycore.jl_b200/tlas.py) in tests/.  See INTEGRATION.md.
#
#   CuTLAS        <: Raycore.AbstractAccel          (src/Raycore.jl:14-49)   replaces Raycore.TLAS   (src/instanced-bvh.jl:261)
#   CuStaticTLAS  <: Raycore.AbstractAdaptedAccel                            replaces Raycore.StaticTLAS (:155)
module RaycoreCUDA

using Raycore, GeometryBasics, StaticArrays, Adapt
import Raycore: Mat3x4f, mat4_to_mat3x4, mat3x4_inverse, TLASHandle, RTRay, RTHitResult, Triangle, Bounds3,
                closest_hit, any_hit, sync!, world_bound, n_instances, n_geometries, wait_for_gpu!, is_valid,
                get_instance, get_instances, update_transform!, update_transforms!, update!, free!

const lib = get(ENV, "RAYCORE_CUDA_LIB", "libraycore_cuda")

const RC_RAYS_ON_DEVICE = UInt32(0x1); const RC_HITS_ON_DEVICE = UInt32(0x2); const RC_MODE_REFERENCE_ORDER = UInt32(0x4)
const RC_BUILD_KEEP_BVH2 = UInt32(0x80); const RC_BUILD_ALLOW_REFIT = UInt32(0x100); const RC_UPDATE_REFIT = UInt32(0x200)
const RC_MODE_WATERTIGHT = UInt32(0x400)

struct RcError <: Exception; code::Int32; msg::String; end
function check(ctx, rc::Int32)
    rc == 0 && return nothing
    msg = unsafe_string(ccall((:rc_last_error, lib), Cstring, (Ptr{Cvoid},), ctx))
    # the reference raises ErrorException via error(...) for handle misuse (src/instanced-bvh.jl:715-718,756-759)
    rc in (1, 2, 3, 4) ? error(msg) : throw(RcError(rc, msg))
end

mutable struct CuStaticTLAS{T} <: Raycore.AbstractAdaptedAccel
    owner::Any          # the CuTLAS; identity of this object is kept across refits, replaced on rebuilds
    generation::Int
end

"What triangle_of needs from a pushed mesh, decomposed once at push! time (not per hit)."
struct MeshSource
    verts::Vector{Point3f}; norms::Vector{Raycore.Normal3f}; uvs::Vector{Point2f}; indices::Vector{UInt32}
end
function MeshSource(nmesh)
    fs = decompose(TriangleFace{UInt32}, nmesh); uvs_raw = GeometryBasics.decompose_uv(nmesh)
    MeshSource(decompose(Point3f, nmesh), Raycore.Normal3f.(decompose_normals(nmesh)), isnothing(uvs_raw) ? Point2f[] : Point2f.(uvs_raw), collect(reinterpret(UInt32, fs)))
end

mutable struct CuTLAS <: Raycore.AbstractAccel
    ctx::Ptr{Cvoid}
    meshes::Dict{UInt32, MeshSource}       # handle id => decomposed source mesh (input order)
    prim_to_face::Dict{UInt32, Vector{UInt32}}
    inst_handles::Vector{UInt32}           # instance position => handle id, refreshed by sync!
    inst_blas::Vector{UInt32}              # instance position => blas_index, refreshed by sync!
    static_tlas::Union{Nothing, CuStaticTLAS}
    generation::Int
    # keep_bvh2 / allow_refit: RC_BUILD_KEEP_BVH2 (reference-order mode, BVH2 read-backs) / RC_BUILD_ALLOW_REFIT (update!(...; refit = true))
    function CuTLAS(device::Integer = -1; keep_bvh2::Bool = false, allow_refit::Bool = false)
        ref = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:rc_create, lib), Int32, (Int32, Ref{Ptr{Cvoid}}), device, ref)
        rc == 0 || error(unsafe_string(ccall((:rc_last_error, lib), Cstring, (Ptr{Cvoid},), C_NULL)))
        t = new(ref[], Dict(), Dict(), UInt32[], UInt32[], nothing, 0)
        flags = (keep_bvh2 ? RC_BUILD_KEEP_BVH2 : UInt32(0)) | (allow_refit ? RC_BUILD_ALLOW_REFIT : UInt32(0))
        flags != 0 && check(t.ctx, ccall((:rc_set_build_flags, lib), Int32, (Ptr{Cvoid}, UInt32), t.ctx, flags))
        finalizer(free!, t)                                    # finalizer(free!, tlas), src/instanced-bvh.jl:355
        t
    end
end

# ---- the reference's convenience constructors (src/instanced-bvh.jl:2276-2324, 2361-2378) ----------------------------------
"""
    CuTLAS(primitives, metadata_fn; device = -1) -> CuStaticTLAS      (TLAS(primitives, metadata_fn), :2276-2324)

One BLAS per primitive with a single identity instance, `Triangle.metadata = metadata_fn(mesh_idx, face_idx)` (face index counted before the
degenerate filter, :2303) and `InstanceDescriptor.instance_id = mesh_idx` (:2316).  Like the reference's, the result is the traversable
(adapted) structure.  The hit record carries 32 bits of metadata: `metadata_fn` must return a 4-byte isbits value (UInt32, Int32, Float32, ...).
"""
function CuTLAS(primitives::AbstractVector, metadata_fn::Function; device::Integer = -1)
    sizeof(typeof(metadata_fn(1, 1))) == 4 || error("metadata_fn must return a 4-byte isbits value (the hit record carries 32 bits of metadata)")
    t = CuTLAS(device)
    for (mi, prim) in enumerate(primitives)
        gb_mesh = prim isa GeometryBasics.Mesh ? prim : GeometryBasics.uv_normal_mesh(prim)      # :2291
        v, _, nmesh = soup(gb_mesh)
        meta = UInt32[reinterpret(UInt32, metadata_fn(mi, i)) for i in 1:size(v, 2)]
        push_soup!(t, v, meta, nmesh, [Mat4f(I)]; instance_ids = UInt32[mi])
    end
    return Adapt.adapt(nothing, t)
end
"""
    CuTLAS(meshes::AbstractVector{<:GeometryBasics.Mesh}; device = -1) -> (CuTLAS, Vector{TLASHandle})      (TLAS(meshes), :2361-2378)
"""
function CuTLAS(meshes::AbstractVector{<:GeometryBasics.Mesh}; device::Integer = -1, kw...)
    isempty(meshes) && error("Cannot create TLAS from empty mesh list")                           # :2362
    t = CuTLAS(device; kw...)
    handles = TLASHandle[push!(t, m) for m in meshes]
    sync!(t)
    return t, handles
end
Base.eltype(::CuTLAS) = Triangle{UInt32}                                                          # :2335-2338
Base.eltype(::CuStaticTLAS{T}) where {T} = T                                                      # :2340-2342

function free!(t::CuTLAS)
    t.ctx == C_NULL && return nothing
    ccall((:rc_destroy, lib), Int32, (Ptr{Cvoid},), t.ctx); t.ctx = C_NULL; nothing
end

# ---- decomposition identical to build_and_append_blas! (src/instanced-bvh.jl:581-600), minus the filter (done on the GPU)
function soup(mesh::GeometryBasics.Mesh)
    nmesh = GeometryBasics.expand_faceviews(mesh)
    fs = decompose(TriangleFace{UInt32}, nmesh); verts = decompose(Point3f, nmesh)
    v = Matrix{Float32}(undef, 9, length(fs))
    for (i, f) in enumerate(fs), k in 1:3, c in 1:3
        v[3 * (k - 1) + c, i] = verts[f[k]][c]
    end
    meta = hasproperty(nmesh, :face_meta) ? UInt32[nmesh.face_meta[f[1]] for f in fs] : nothing
    return v, meta, nmesh
end

function Base.push!(t::CuTLAS, mesh::GeometryBasics.Mesh, transforms::AbstractVector{Mat4f};
                    instance_ids::Union{Nothing, AbstractVector{<:Integer}} = nothing, sbt_offset::UInt32 = UInt32(0))
    instance_ids !== nothing && length(instance_ids) != length(transforms) &&
        throw(ArgumentError("instance_ids length $(length(instance_ids)) != transforms length $(length(transforms))"))
    v, meta, nmesh = soup(mesh)
    return push_soup!(t, v, meta, nmesh, transforms; instance_ids)
end
function push_soup!(t::CuTLAS, v::Matrix{Float32}, meta, nmesh, transforms::AbstractVector{Mat4f}; instance_ids = nothing)
    xf = [mat4_to_mat3x4(m) for m in transforms]
    inv = [mat3x4_inverse(m) for m in xf]          # computed with Raycore's own code => bit-identical descriptors
    ids = instance_ids === nothing ? C_NULL : UInt32.(instance_ids)
    h = Ref{UInt32}(0)
    check(t.ctx, ccall((:rc_push, lib), Int32,
        (Ptr{Cvoid}, Ptr{Float32}, UInt32, Ptr{UInt32}, Ptr{Float32}, Ptr{Float32}, Ptr{UInt32}, UInt32, UInt32, Ref{UInt32}),
        t.ctx, v, size(v, 2), meta === nothing ? C_NULL : meta, reinterpret(Float32, xf), reinterpret(Float32, inv), ids,
        length(xf), 0, h))
    t.meshes[h[]] = MeshSource(nmesh)
    return TLASHandle(h[])
end
Base.push!(t::CuTLAS, mesh::GeometryBasics.Mesh, transform::Mat4f = Mat4f(I); instance_id::UInt32 = UInt32(0), sbt_offset::UInt32 = UInt32(0)) =
    push!(t, mesh, [transform]; instance_ids = [instance_id])

# ---- serialised geometry (to_gpu(ArrayType, blas::BLAS), src/kernel-abstractions.jl:31-36: upload a built BLAS instead of rebuilding) ----
"Bytes of the handle's built geometry (BVH2 in BVHNode2 order, wide nodes, sorted triangles, hull, normals): write them to disk, restore with `push_exported!`."
function export_geometry(t::CuTLAS, h::TLASHandle)
    n = Ref{UInt64}(0)
    check(t.ctx, ccall((:rc_export_geometry, lib), Int32, (Ptr{Cvoid}, UInt32, Ptr{Cvoid}, UInt64, Ref{UInt64}), t.ctx, h.id, C_NULL, 0, n))
    blob = Vector{UInt8}(undef, n[])
    check(t.ctx, ccall((:rc_export_geometry, lib), Int32, (Ptr{Cvoid}, UInt32, Ptr{Cvoid}, UInt64, Ref{UInt64}), t.ctx, h.id, blob, length(blob), n))
    return blob
end
"Host-side vetting of `export_geometry` bytes (header, layout version, size, payload hash) without a context or a GPU; errors like the import would."
function check_exported(blob::Vector{UInt8})
    n = Ref{UInt32}(0); f = Ref{UInt32}(0); hn = Ref{UInt32}(0)
    rc = ccall((:rc_check_exported, lib), Int32, (Ptr{Cvoid}, UInt64, Ref{UInt32}, Ref{UInt32}, Ref{UInt32}), blob, length(blob), n, f, hn)
    rc == 0 || error(unsafe_string(ccall((:rc_last_error, lib), Cstring, (Ptr{Cvoid},), C_NULL)))
    return (n_triangles = Int(n[]), n_faces = Int(f[]), has_normals = hn[] != 0)
end
"push! of a geometry restored from `export_geometry` bytes (no builder kernel runs).  `mesh` is the caller's copy used to materialise Triangles."
function push_exported!(t::CuTLAS, blob::Vector{UInt8}, mesh::GeometryBasics.Mesh, transforms::AbstractVector{Mat4f};
                        instance_ids::Union{Nothing, AbstractVector{<:Integer}} = nothing)
    instance_ids !== nothing && length(instance_ids) != length(transforms) &&
        throw(ArgumentError("instance_ids length $(length(instance_ids)) != transforms length $(length(transforms))"))
    xf = [mat4_to_mat3x4(m) for m in transforms]; inv = [mat3x4_inverse(m) for m in xf]
    ids = instance_ids === nothing ? C_NULL : UInt32.(instance_ids)
    h = Ref{UInt32}(0)
    check(t.ctx, ccall((:rc_push_exported, lib), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, UInt64, Ptr{Float32}, Ptr{Float32}, Ptr{UInt32}, UInt32, Ref{UInt32}),
        t.ctx, blob, length(blob), reinterpret(Float32, xf), reinterpret(Float32, inv), ids, length(xf), h))
    t.meshes[h[]] = MeshSource(GeometryBasics.expand_faceviews(mesh))
    return TLASHandle(h[])
end

function Base.delete!(t::CuTLAS, h::TLASHandle)::Bool
    d = Ref{Int32}(0); check(t.ctx, ccall((:rc_delete, lib), Int32, (Ptr{Cvoid}, UInt32, Ref{Int32}), t.ctx, h.id, d)); d[] != 0
end

function update_transforms!(t::CuTLAS, h::TLASHandle, transforms::AbstractVector{Mat3x4f})
    inv = [mat3x4_inverse(m) for m in transforms]
    check(t.ctx, ccall((:rc_update_transforms, lib), Int32, (Ptr{Cvoid}, UInt32, Ptr{Float32}, Ptr{Float32}, UInt32),
        t.ctx, h.id, reinterpret(Float32, collect(transforms)), reinterpret(Float32, inv), length(transforms)))
end
update_transforms!(t::CuTLAS, h::TLASHandle, ts::AbstractVector{Mat4f}) = update_transforms!(t, h, map(mat4_to_mat3x4, ts))
function update_transform!(t::CuTLAS, h::TLASHandle, m::Union{Mat4f, Mat3x4f})
    n = ccall((:rc_n_instances_of, lib), UInt32, (Ptr{Cvoid}, UInt32), t.ctx, h.id)
    is_valid(t, h) && n != 1 && error("Handle has $n instances, use update_transforms! for multiple")
    update_transforms!(t, h, [m isa Mat4f ? mat4_to_mat3x4(m) : m])
end
"update!(tlas, handle, mesh) (:808-857).  `refit = true`: re-fit the kept radix tree when only vertices moved (needs `allow_refit`); returns whether the library re-fitted."
function update!(t::CuTLAS, h::TLASHandle, mesh; refit::Bool = false)
    v, meta, nmesh = soup(mesh)
    check(t.ctx, ccall((:rc_update_geometry, lib), Int32, (Ptr{Cvoid}, UInt32, Ptr{Float32}, UInt32, Ptr{UInt32}, UInt32),
        t.ctx, h.id, v, size(v, 2), meta === nothing ? C_NULL : meta, refit ? RC_UPDATE_REFIT : UInt32(0)))
    t.meshes[h.id] = MeshSource(nmesh)
    return ccall((:rc_last_update_refitted, lib), Int32, (Ptr{Cvoid},), t.ctx) != 0
end

function sync!(t::CuTLAS)
    a = Ref{Int32}(0); check(t.ctx, ccall((:rc_sync, lib), Int32, (Ptr{Cvoid}, Ref{Int32}), t.ctx, a))
    if a[] == 2 || t.static_tlas === nothing           # rebuild => new adapted object; refit keeps identity (test_mesh_update.jl:214)
        t.generation += 1; t.static_tlas = CuStaticTLAS{Triangle{UInt32}}(t, t.generation); empty!(t.prim_to_face)
        # instance position => (handle, BLAS) tables for triangle_of: read once per rebuild, not per hit (compaction renumbers both)
        n = ccall((:rc_n_total_instances, lib), UInt32, (Ptr{Cvoid},), t.ctx)
        t.inst_handles = Vector{UInt32}(undef, n); t.inst_blas = Vector{UInt32}(undef, n)
        ccall((:rc_get_instance_handles, lib), Int32, (Ptr{Cvoid}, Ptr{UInt32}, UInt32), t.ctx, t.inst_handles, n)
        for hid in unique(t.inst_handles)
            d = get_instances(t, TLASHandle(hid))
            for (k, p) in enumerate(findall(==(hid), t.inst_handles)); t.inst_blas[p] = d[k].blas_index; end
        end
    end
    return t
end
Adapt.adapt_structure(to, t::CuTLAS) = (sync!(t); t.static_tlas)          # src/instanced-bvh.jl:1085-1102

is_valid(t::CuTLAS, h::TLASHandle) = ccall((:rc_is_valid, lib), Int32, (Ptr{Cvoid}, UInt32), t.ctx, h.id) != 0
n_instances(t::CuTLAS) = Int(ccall((:rc_n_instances, lib), UInt32, (Ptr{Cvoid},), t.ctx))
n_instances(t::CuTLAS, h::TLASHandle) = Int(ccall((:rc_n_instances_of, lib), UInt32, (Ptr{Cvoid}, UInt32), t.ctx, h.id))
n_geometries(t::CuTLAS) = Int(ccall((:rc_n_geometries, lib), UInt32, (Ptr{Cvoid},), t.ctx))
function world_bound(t::CuTLAS)
    b = Vector{Float32}(undef, 6); check(t.ctx, ccall((:rc_world_bound, lib), Int32, (Ptr{Cvoid}, Ptr{Float32}), t.ctx, b))
    Bounds3(Point3f(b[1:3]...), Point3f(b[4:6]...))
end
wait_for_gpu!(t::CuTLAS) = (check(t.ctx, ccall((:rc_wait, lib), Int32, (Ptr{Cvoid},), t.ctx)); t)
function get_instances(t::CuTLAS, h::TLASHandle)
    out = Vector{Raycore.InstanceDescriptor}(undef, max(1, n_instances(t, h)))
    check(t.ctx, ccall((:rc_get_instances, lib), Int32, (Ptr{Cvoid}, UInt32, Ptr{Cvoid}), t.ctx, h.id, out)); out
end
get_instance(t::CuTLAS, h::TLASHandle, i::Integer = 1) = get_instances(t, h)[i]

# ---- queries -------------------------------------------------------------------------------------------------------
"Batched entry, same shape as Lava.trace_closest_hits!(hits, rays, accel, n) (docs/src/hw_acceleration.md:143-146)."
function trace_closest_hits!(hits::Vector{RTHitResult}, rays::Vector{RTRay}, s::CuStaticTLAS, n::Integer = length(rays); any = false, flags = UInt32(0))
    s.owner.static_tlas === s || error("stale CuStaticTLAS: re-adapt per dispatch (src/instanced-bvh.jl:221-226)")
    f = any ? :rc_trace_any : :rc_trace_closest
    check(s.owner.ctx, ccall((f, lib), Int32, (Ptr{Cvoid}, Ptr{RTRay}, Ptr{RTHitResult}, UInt64, UInt32), s.owner.ctx, rays, hits, n, flags))
    hits
end
Raycore.trace_rays(s::CuStaticTLAS, rays::AbstractVector{<:Raycore.AbstractRay}) =
    trace_closest_hits!(Vector{RTHitResult}(undef, length(rays)), [RTRay(r.o..., r.t_min, r.d..., r.t_max) for r in rays], s)

function hit_tuple(s::CuStaticTLAS, h::RTHitResult, tri_of)
    h.hit == 0 && return (false, Raycore.empty_triangle(Triangle{UInt32}), 0f0, SVector{3, Float32}(0, 0, 0), UInt32(0))   # :2019-2022
    w = 1f0 - h.bary_u - h.bary_v
    (true, tri_of(h), h.t, SVector{3, Float32}(w, h.bary_u, h.bary_v), h.instance_id + UInt32(1))                              # :2010-2017
end
# Per-ray methods: host-side convenience (n = 1 batch).  Device-side per-ray calls from user KA kernels are replaced by the
# batched entry above — the library owns traversal; see INTEGRATION.md "What changes for callers".
closest_hit(s::CuStaticTLAS, ray::Raycore.AbstractRay) =
    hit_tuple(s, trace_closest_hits!([RTHitResult(0, 0, 0, 0, 0, 0, 0, 0)], [RTRay(ray.o..., ray.t_min, ray.d..., ray.t_max)], s)[1], h -> triangle_of(s.owner, h))
any_hit(s::CuStaticTLAS, ray::Raycore.AbstractRay) =
    hit_tuple(s, trace_closest_hits!([RTHitResult(0, 0, 0, 0, 0, 0, 0, 0)], [RTRay(ray.o..., 0f0, ray.d..., ray.t_max)], s; any = true)[1], h -> triangle_of(s.owner, h))

"Materialise Raycore's Triangle (vertices, normals, uv, metadata) for a hit from the caller-side copy of the mesh: two table look-ups and one build_triangle per hit (the mesh was decomposed at push! time, the instance tables at sync!)."
function triangle_of(t::CuTLAS, h::RTHitResult)
    hid = t.inst_handles[h.instance_id + 1]
    blas = t.inst_blas[h.instance_id + 1]
    faces = get!(t.prim_to_face, blas) do
        n = ccall((:rc_blas_n_prims, lib), UInt32, (Ptr{Cvoid}, UInt32), t.ctx, blas); out = Vector{UInt32}(undef, n)
        ccall((:rc_read_blas_faces, lib), Int32, (Ptr{Cvoid}, UInt32, Ptr{UInt32}, UInt32), t.ctx, blas, out, n); out
    end
    m = t.meshes[hid]
    Raycore.build_triangle(m.verts, m.norms, m.uvs, m.indices, faces[h.primitive_id + 1] + 1, h._pad2)      # _pad2 carries Triangle.metadata
end

# ---- BLAS4 / build_blas4 / closest_hit4 / any_hit4 (src/bvh4.jl:154-163, 511-523, 606-766) --------------------------------
"One geometry on its own 4-wide BVH (rc_blas4_*): the library's wide BVH under an identity instance."
mutable struct CuBLAS4
    ptr::Ptr{Cvoid}
    primitives::Vector            # the caller's Triangles, in input order (closest_hit4 returns one of them)
    prim_to_input::Vector{UInt32} # primitive_id (position after the degenerate filter) => index into `primitives` (0-based)
end
"build_blas4(primitives) (:511-523) on the GPU; `primitives::AbstractVector{<:Triangle}` as in the reference."
function build_blas4_cuda(primitives::AbstractVector{<:Triangle}; device::Integer = -1)
    isempty(primitives) && error("Cannot build BLAS4 from empty primitive list")                    # :513
    v = Matrix{Float32}(undef, 9, length(primitives))
    for (i, tri) in enumerate(primitives), k in 1:3, c in 1:3; v[3 * (k - 1) + c, i] = tri.vertices[k][c]; end
    meta = UInt32[reinterpret(UInt32, tri.metadata) for tri in primitives]
    ref = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:rc_blas4_build, lib), Int32, (Int32, Ptr{Float32}, UInt32, Ptr{UInt32}, UInt32, Ref{Ptr{Cvoid}}), device, v, size(v, 2), meta, 0, ref)
    rc == 0 || error(unsafe_string(ccall((:rc_blas4_last_error, lib), Cstring, (Ptr{Cvoid},), C_NULL)))
    n = Ref{UInt32}(0); ccall((:rc_blas4_info, lib), Int32, (Ptr{Cvoid}, Ref{UInt32}, Ptr{UInt32}, Ptr{Float32}), ref[], n, C_NULL, C_NULL)
    faces = Vector{UInt32}(undef, n[])
    ccall((:rc_read_blas_faces, lib), Int32, (Ptr{Cvoid}, UInt32, Ptr{UInt32}, UInt32), ccall((:rc_blas4_context, lib), Ptr{Cvoid}, (Ptr{Cvoid},), ref[]), 1, faces, n[])
    b = CuBLAS4(ref[], collect(primitives), faces)
    finalizer(x -> (x.ptr != C_NULL && ccall((:rc_blas4_destroy, lib), Int32, (Ptr{Cvoid},), x.ptr); x.ptr = C_NULL), b)
    b
end
function trace4!(hits::Vector{RTHitResult}, rays::Vector{RTRay}, b::CuBLAS4; any = false, flags = UInt32(0))
    f = any ? :rc_blas4_trace_any : :rc_blas4_trace_closest
    rc = ccall((f, lib), Int32, (Ptr{Cvoid}, Ptr{RTRay}, Ptr{RTHitResult}, UInt64, UInt32), b.ptr, rays, hits, length(rays), flags)
    rc == 0 || error(unsafe_string(ccall((:rc_blas4_last_error, lib), Cstring, (Ptr{Cvoid},), b.ptr)))
    hits
end
function hit_tuple4(b::CuBLAS4, h::RTHitResult)                                                         # (hit, primitive, distance, barycentric), :606
    h.hit == 0 && return (false, Raycore.empty_triangle(eltype(b.primitives)), 0f0, SVector{3, Float32}(0, 0, 0))
    (true, b.primitives[b.prim_to_input[h.primitive_id + 1] + 1], h.t, SVector{3, Float32}(1f0 - h.bary_u - h.bary_v, h.bary_u, h.bary_v))
end
Raycore.closest_hit4(b::CuBLAS4, ray::Raycore.AbstractRay) =
    hit_tuple4(b, trace4!([RTHitResult(0, 0, 0, 0, 0, 0, 0, 0)], [RTRay(ray.o..., ray.t_min, ray.d..., ray.t_max)], b)[1])
Raycore.any_hit4(b::CuBLAS4, ray::Raycore.AbstractRay) =
    hit_tuple4(b, trace4!([RTHitResult(0, 0, 0, 0, 0, 0, 0, 0)], [RTRay(ray.o..., 0f0, ray.d..., ray.t_max)], b; any = true)[1])

# ---- several GPUs behind one handle (rc_multi_*): replicated scene, sharded queries, one process ------------------------------
mutable struct CuMultiTLAS
    ptr::Ptr{Cvoid}
end
function CuMultiTLAS(devices::Union{Nothing, AbstractVector{<:Integer}} = nothing)
    ref = Ref{Ptr{Cvoid}}(C_NULL); dv = devices === nothing ? Int32[] : Int32.(devices)
    rc = ccall((:rc_multi_create, lib), Int32, (Ptr{Int32}, UInt32, Ref{Ptr{Cvoid}}), isempty(dv) ? C_NULL : dv, length(dv), ref)
    rc == 0 || error(unsafe_string(ccall((:rc_multi_last_error, lib), Cstring, (Ptr{Cvoid},), C_NULL)))
    m = CuMultiTLAS(ref[]); finalizer(x -> (x.ptr != C_NULL && ccall((:rc_multi_destroy, lib), Int32, (Ptr{Cvoid},), x.ptr); x.ptr = C_NULL), m); m
end
mcheck(m::CuMultiTLAS, rc::Int32) = rc == 0 ? nothing : error(unsafe_string(ccall((:rc_multi_last_error, lib), Cstring, (Ptr{Cvoid},), m.ptr)))
function Base.push!(m::CuMultiTLAS, mesh::GeometryBasics.Mesh, transforms::AbstractVector{Mat4f}; instance_ids = nothing)
    v, meta, _ = soup(mesh); xf = [mat4_to_mat3x4(x) for x in transforms]; inv = [mat3x4_inverse(x) for x in xf]; h = Ref{UInt32}(0)
    mcheck(m, ccall((:rc_multi_push, lib), Int32, (Ptr{Cvoid}, Ptr{Float32}, UInt32, Ptr{UInt32}, Ptr{Float32}, Ptr{Float32}, Ptr{UInt32}, UInt32, UInt32, Ref{UInt32}),
        m.ptr, v, size(v, 2), meta === nothing ? C_NULL : meta, reinterpret(Float32, xf), reinterpret(Float32, inv), instance_ids === nothing ? C_NULL : UInt32.(instance_ids), length(xf), 0, h))
    TLASHandle(h[])
end
sync!(m::CuMultiTLAS) = (mcheck(m, ccall((:rc_multi_sync, lib), Int32, (Ptr{Cvoid}, Ptr{Int32}), m.ptr, C_NULL)); m)
"closest hits of `rays`, sharded over every device of the handle; identical to the single-device result"
trace_closest_hits!(hits::Vector{RTHitResult}, rays::Vector{RTRay}, m::CuMultiTLAS; any = false, flags = UInt32(0)) =
    (mcheck(m, ccall((any ? :rc_multi_trace_any : :rc_multi_trace_closest, lib), Int32, (Ptr{Cvoid}, Ptr{RTRay}, Ptr{RTHitResult}, UInt64, UInt32), m.ptr, rays, hits, length(rays), flags)); hits)
function Raycore.view_factors(m::CuMultiTLAS; rays_per_triangle = 10000, seed = 0)
    n = Ref{UInt32}(0); ccall((:rc_sizes, lib), Int32, (Ptr{Cvoid}, Ptr{UInt32}, Ptr{UInt32}, Ref{UInt32}, Ptr{UInt32}), ccall((:rc_multi_context, lib), Ptr{Cvoid}, (Ptr{Cvoid}, UInt32), m.ptr, 0), C_NULL, C_NULL, n, C_NULL)
    out = zeros(UInt32, n[], n[]); sk = Ref{UInt64}(0)
    mcheck(m, ccall((:rc_multi_view_factors, lib), Int32, (Ptr{Cvoid}, UInt32, UInt64, Ptr{UInt32}, Ref{UInt64}), m.ptr, rays_per_triangle, seed, out, sk))
    permutedims(out)
end

# ---- analysis (src/kernels.jl) -----------------------------------------------------------------------------------------
function Raycore.get_illumination(t::CuTLAS, viewdir; grid_size = 1000)
    sync!(t); n = Ref{UInt32}(0); ccall((:rc_sizes, lib), Int32, (Ptr{Cvoid}, Ptr{UInt32}, Ptr{UInt32}, Ref{UInt32}, Ptr{UInt32}), t.ctx, C_NULL, C_NULL, n, C_NULL)
    out = zeros(Float32, n[]); check(t.ctx, ccall((:rc_get_illumination, lib), Int32, (Ptr{Cvoid}, Ptr{Float32}, UInt32, Ptr{Float32}, UInt32), t.ctx, Float32[viewdir...], grid_size, out, n[])); out
end
function Raycore.get_centroid(t::CuTLAS, viewdir; grid_size = 32)
    sync!(t); c = zeros(Float32, 3); nh = Ref{UInt32}(0); pts = Matrix{Float32}(undef, 3, grid_size^2)
    check(t.ctx, ccall((:rc_get_centroid, lib), Int32, (Ptr{Cvoid}, Ptr{Float32}, UInt32, Ptr{Float32}, Ref{UInt32}, Ptr{Float32}), t.ctx, Float32[viewdir...], grid_size, c, nh, pts))
    [Point3f(pts[:, i]...) for i in 1:nh[]], Point3f(c...)
end
function Raycore.view_factors(t::CuTLAS; rays_per_triangle = 10000, seed = 0)
    sync!(t); n = Ref{UInt32}(0); ccall((:rc_sizes, lib), Int32, (Ptr{Cvoid}, Ptr{UInt32}, Ptr{UInt32}, Ref{UInt32}, Ptr{UInt32}), t.ctx, C_NULL, C_NULL, n, C_NULL)
    out = zeros(UInt32, n[], n[]); sk = Ref{UInt64}(0)   # C side is row-major [src][hit] = Julia's transpose
    check(t.ctx, ccall((:rc_view_factors, lib), Int32, (Ptr{Cvoid}, UInt32, UInt64, Ptr{UInt32}, UInt32, UInt32, UInt32, Ref{UInt64}), t.ctx, rays_per_triangle, seed, out, 0, n[], 0, sk))
    permutedims(out)                                       # result[src_meta, hit_meta]
end

# ---- collision broad phase (src/collision.jl:189-261) ---------------------------------------------------------------------
function Raycore.collide_instances(t::CuTLAS)
    sync!(t); n = Ref{UInt64}(0)
    check(t.ctx, ccall((:rc_collide_instances, lib), Int32, (Ptr{Cvoid}, Ptr{Raycore.ContactPair}, UInt64, Ref{UInt64}), t.ctx, C_NULL, 0, n))
    pairs = Vector{Raycore.ContactPair}(undef, n[])
    n[] > 0 && check(t.ctx, ccall((:rc_collide_instances, lib), Int32, (Ptr{Cvoid}, Ptr{Raycore.ContactPair}, UInt64, Ref{UInt64}), t.ctx, pairs, n[], n))
    pairs
end
function Raycore.collide_instances_any(t::CuTLAS, a::TLASHandle, b::TLASHandle)
    sync!(t); r = Ref{Int32}(0)
    check(t.ctx, ccall((:rc_collide_instances_any, lib), Int32, (Ptr{Cvoid}, UInt32, UInt32, Ref{Int32}), t.ctx, a.id, b.id, r)); r[] != 0
end

# ---- wavefront stages on device-resident queues (docs/src/wavefront-renderer.jl:185-362) -----------------------------------
# Each function replaces one `kernel!(backend)(...; ndrange)` launch of the reference renderer; queues are CuArray{RTRay},
# CuArray{RTHitResult} and CuArray{UInt8} (CuPtr arguments), lights a host Vector{Point3f}.
const RC_ON_DEVICE = RC_RAYS_ON_DEVICE | RC_HITS_ON_DEVICE
# device arrays come from CUDA.jl (pointer(::CuArray) is a CuPtr); the C ABI takes plain addresses
devptr(a) = reinterpret(Ptr{Cvoid}, pointer(a))
set_normals!(t::CuTLAS, h::TLASHandle, normals9::Matrix{Float32}) =       # 9 x n_faces: Triangle.normals per submitted face
    check(t.ctx, ccall((:rc_set_normals, lib), Int32, (Ptr{Cvoid}, UInt32, Ptr{Float32}, UInt32, UInt32), t.ctx, h.id, normals9, size(normals9, 2), 0))
generate_primary_rays!(t::CuTLAS, rays, width, height, camera_pos, focal_length, aspect; nsamples = 1, seed = 0) =
    check(t.ctx, ccall((:rc_generate_primary_rays, lib), Int32, (Ptr{Cvoid}, UInt32, UInt32, UInt32, Ptr{Float32}, Float32, Float32, UInt64, Ptr{Cvoid}, UInt32),
                       t.ctx, width, height, nsamples, Float32[camera_pos...], focal_length, aspect, seed, devptr(rays), 0))
intersect_primary_rays!(t::CuTLAS, rays, hits) =
    check(t.ctx, ccall((:rc_trace_closest, lib), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, UInt64, UInt32), t.ctx, devptr(rays), devptr(hits), length(rays), RC_ON_DEVICE))
generate_shadow_rays!(t::CuTLAS, rays, hits, lights::Vector{Point3f}, shadow_rays; shadow_bias = 0.01f0) =
    check(t.ctx, ccall((:rc_generate_shadow_rays, lib), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, UInt64, Ptr{Float32}, UInt32, Float32, Ptr{Cvoid}, UInt32),
                       t.ctx, devptr(rays), devptr(hits), length(rays), reinterpret(Float32, lights), length(lights), shadow_bias, devptr(shadow_rays), 0))
test_shadow_rays!(t::CuTLAS, shadow_rays, visible) =
    check(t.ctx, ccall((:rc_test_shadow_rays, lib), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, UInt64, Ptr{Cvoid}, UInt32), t.ctx, devptr(shadow_rays), length(shadow_rays), devptr(visible), 0))
"stages 3 + 4 in one kernel: no shadow-ray queue"
shadow_visibility!(t::CuTLAS, rays, hits, lights::Vector{Point3f}, visible; shadow_bias = 0.01f0) =
    check(t.ctx, ccall((:rc_shadow_visibility, lib), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, UInt64, Ptr{Float32}, UInt32, Float32, Ptr{Cvoid}, UInt32),
                       t.ctx, devptr(rays), devptr(hits), length(rays), reinterpret(Float32, lights), length(lights), shadow_bias, devptr(visible), 0))

end # module