#!/usr/bin/env julia
# make_ref_golden.jl INPUT_DIR OUTPUT_DIR — golden vectors from the REAL Raycore.jl, for a maintainer who has Julia.
#
# The build image of this repository has no Julia, so tests/golden/*.npz freeze the C oracle (oracle/oracle.c), which is pinned to the
# reference only through the reference's own known-answer tests (tests/kat.py).  This script closes that gap where Julia is available:
#
#     python tests/golden/make_golden.py --export-inputs /tmp/rc_inputs          # scenes + rays of the committed fixtures, as .npz
#     julia --project=<env with Raycore, GeometryBasics, StaticArrays, NPZ> tests/golden/make_ref_golden.jl /tmp/rc_inputs tests/golden
#     python -m pytest tests/test_golden.py -k reference_fixture                 # oracle == Raycore.jl, bit for bit
#
# It builds every scene with the reference's own build_blas / build_tlas (src/instanced-bvh.jl:1376-1443, :1605-1640) from the same
# triangles, transforms and instance ids, traces the same rays with closest_hit / any_hit (:1902-2140) on the CPU, and writes
# tests/golden/ref_<scene>.npz.  Raycore returns the hit triangle rather than a primitive index, so the fixture stores the triangle's
# vertices and metadata; tests/test_golden.py matches them against the oracle's triangle at its reported primitive_id.
using Raycore, GeometryBasics, StaticArrays, LinearAlgebra, NPZ
import Raycore: Mat3x4f, mat3x4_inverse, is_degenerate, InstanceDescriptor, build_blas, build_tlas, closest_hit, any_hit, Ray, Normal3f

const RTriangle = Raycore.Triangle   # (GeometryBasics has a Triangle too)

"Triangles of one push: rows of `verts` (n x 9) in input order, degenerate faces dropped as build_and_append_blas! does (:591-600); metadata = face_meta[i] or the 1-based face index (:595)."
function triangles_of(verts::AbstractMatrix{Float32}, meta)
    tris = RTriangle{UInt32}[]
    for i in 1:size(verts, 1)
        vs = SVector(Point3f(verts[i, 1:3]...), Point3f(verts[i, 4:6]...), Point3f(verts[i, 7:9]...))
        is_degenerate(vs) && continue
        n = Normal3f(0, 0, 1)
        push!(tris, RTriangle(vs, SVector(n, n, n), SVector(Vec3f(0), Vec3f(0), Vec3f(0)), SVector(Point2f(0, 0), Point2f(1, 0), Point2f(1, 1)),
                              meta === nothing ? UInt32(i) : UInt32(meta[i])))
    end
    tris
end

function record!(out, k, res)
    hit, tri, t, bary, inst = res
    out["hit"][k] = hit ? 0x01 : 0x00
    out["t"][k] = t; out["bary_u"][k] = bary[2]; out["bary_v"][k] = bary[3]
    out["instance_id"][k] = hit ? UInt32(inst - 1) : UInt32(0)          # 0-based instance position, as RTHitResult.instance_id
    out["meta"][k] = hit ? UInt32(tri.metadata) : UInt32(0)
    if hit
        for c in 1:3, a in 1:3; out["verts"][k, 3 * (c - 1) + a] = tri.vertices[c][a]; end
    end
end

function main(indir, outdir)
    for f in sort(filter(endswith("_inputs.npz"), readdir(indir)))
        name = replace(f, "_inputs.npz" => "")
        d = npzread(joinpath(indir, f))
        blases = Raycore.BLAS[]; instances = InstanceDescriptor[]
        for p in 0:Int(d["n_pushes"][1]) - 1
            meta = haskey(d, "meta_$p") ? d["meta_$p"] : nothing
            push!(blases, build_blas(triangles_of(d["verts_$p"], meta)))
            xf = d["xf_$p"]                                                # m x 12, row-major 3x4 == the memory order of Mat3x4f (SMatrix{4,3})
            ids = haskey(d, "ids_$p") ? d["ids_$p"] : zeros(UInt32, size(xf, 1))
            for i in 1:size(xf, 1)
                t = Mat3x4f(xf[i, :]...)
                push!(instances, InstanceDescriptor(UInt32(length(blases)), UInt32(ids[i]), t, mat3x4_inverse(t), UInt32(0)))
            end
        end
        blases = [b for b in blases]                                       # concrete element type for build_tlas
        tlas = build_tlas(blases, instances)
        rays = d["rays"]                                                   # n x 8: origin, t_min, direction, t_max (RTRay)
        n = size(rays, 1)
        mk() = Dict{String, Any}("hit" => zeros(UInt8, n), "t" => zeros(Float32, n), "bary_u" => zeros(Float32, n), "bary_v" => zeros(Float32, n),
                                  "instance_id" => zeros(UInt32, n), "meta" => zeros(UInt32, n), "verts" => zeros(Float32, n, 9))
        cl, an = mk(), mk()
        for k in 1:n
            ray = Ray(o = Point3f(rays[k, 1:3]...), d = Vec3f(rays[k, 5:7]...), t_min = rays[k, 4], t_max = rays[k, 8])
            record!(cl, k, closest_hit(tlas, ray))
            record!(an, k, any_hit(tlas, ray))
        end
        out = Dict{String, Any}("rays" => rays)
        for (k, v) in cl; out["closest_" * k] = v; end
        for (k, v) in an; out["any_" * k] = v; end
        npzwrite(joinpath(outdir, "ref_$name.npz"), out)
        println(name, ": ", n, " rays, hit rate ", sum(cl["hit"]) / n)
    end
end

main(ARGS[1], ARGS[2])
