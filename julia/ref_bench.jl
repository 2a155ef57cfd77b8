# ref_bench.jl -- times the hot path of the UNMODIFIED reference package on the host CPU, for anyone who has Julia.
#
# UNEXECUTED in this repository's environment (no Julia in the image, SURVEY.md 8c): `bench.py --impl reference` times
# the C restatement of the same loops instead.  Usage (from a checkout of cohensbw/ElPhDynamics with its Manifest):
#
#     julia --project=. ref_bench.jl examples/holstein_langevin_square.toml 32 20.0
#
# Prints one JSON line per quantity in the units of bench.py: M^T M matvecs/s, CG iterations/s, Langevin steps/s.
# The reference is single-threaded (src/ElPhDynamics.jl:74-75 pins BLAS and FFTW to one thread).
using ElPhDynamics, Random, Printf, LinearAlgebra
using ElPhDynamics.Models: mulMᵀM!, update_model!
using ElPhDynamics.LangevinDynamics: evolve!
using ElPhDynamics.KPMPreconditioners: setup!

function main(args)
    input = args[1]
    L = length(args) > 1 ? parse(Int, args[2]) : 32
    β = length(args) > 2 ? parse(Float64, args[3]) : 20.0
    # process_input_file builds model, dynamics, Fourier accelerator and preconditioner exactly as simulate() does
    # (src/ProcessInputFile.jl:30-600); the lattice size and β of the shipped example are overridden here
    input_dict = ElPhDynamics.TOML.parsefile(input)
    input_dict["lattice"]["L"] = [L, L, 1]
    input_dict["model"]["beta"] = β
    model, Gr, μ_tuner, sim_params, simulation_dynamics, burnin_dynamics, fa, preconditioner, container =
        ElPhDynamics.process_input_file(input_dict, 1)
    update_model!(model)
    n = model.Ndim
    v, y, x = randn(model.rng, n), zeros(n), zeros(n)
    mulMᵀM!(y, model, v)                                   # warm-up / compilation
    reps = 200
    t = @elapsed for _ in 1:reps; mulMᵀM!(y, model, v); end
    @printf("{\"metric\": \"MTM matvecs/s\", \"value\": %.3f, \"threads\": 1, \"impl\": \"ElPhDynamics.jl\"}\n", reps / t)
    b = similar(v); ElPhDynamics.Models.mulMᵀ!(b, model, v)
    fill!(x, 0.0); ldiv!(x, model, b)                      # warm-up
    fill!(x, 0.0)
    t = @elapsed iters, err, flag = ldiv!(x, model, b)
    @printf("{\"metric\": \"CG iterations/s\", \"value\": %.3f, \"iters\": %d, \"flag\": %d}\n", iters / t, iters, flag)
    evolve!(model, simulation_dynamics, fa, preconditioner) # warm-up
    nsteps = 5
    t = @elapsed for _ in 1:nsteps; evolve!(model, simulation_dynamics, fa, preconditioner); end
    @printf("{\"metric\": \"Langevin steps/s\", \"value\": %.4f}\n", nsteps / t)
end

main(ARGS)
