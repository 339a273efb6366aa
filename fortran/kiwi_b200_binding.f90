! kiwi_b200_binding.f90 -- iso_c_binding interface to libkiwi_b200.so (include/kiwi_b200.h).
!
! UNTESTED SOURCE: the build image has no Fortran compiler (SURVEY.md, fact 2).  It is the stub a
! Kiwi maintainer adds to the reference tree so that minimizer_engine.f90 calls the B200 engine
! for the path  set_source_params -> calculate_seismograms -> scale_seismograms ->
! calculate_misfits  (minimizer_engine.f90:500-523, 885-945).  See INTEGRATION.md.
module kiwi_b200_binding

    use iso_c_binding
    implicit none

    interface

        function kiwi_last_error() bind(C, name="kiwi_last_error") result(msg)
            import :: c_ptr
            type(c_ptr) :: msg
        end function

        function kiwi_create(device) bind(C, name="kiwi_create") result(ctx)
            import :: c_ptr, c_int
            integer(c_int), value :: device
            type(c_ptr) :: ctx
        end function

        subroutine kiwi_destroy(ctx) bind(C, name="kiwi_destroy")
            import :: c_ptr
            type(c_ptr), value :: ctx
        end subroutine

        function kiwi_gfdb_create(nx, nz, ng, dt, dx, dz, firstx, firstz) bind(C, name="kiwi_gfdb_create") result(db)
            import :: c_ptr, c_int, c_float
            integer(c_int), value :: nx, nz, ng
            real(c_float), value :: dt, dx, dz, firstx, firstz
            type(c_ptr) :: db
        end function

        ! one call per stored trace while walking the HDF5 chunks with gfdb_get_trace (gfdb.f90:830-863)
        function kiwi_gfdb_save_array(db, ix, iz, ig, span0, n, data) bind(C, name="kiwi_gfdb_save_array") result(rc)
            import :: c_ptr, c_int, c_float
            type(c_ptr), value :: db
            integer(c_int), value :: ix, iz, ig, span0, n
            real(c_float), intent(in) :: data(*)
            integer(c_int) :: rc
        end function

        function kiwi_set_database(ctx, db) bind(C, name="kiwi_set_database") result(rc)
            import :: c_ptr, c_int
            type(c_ptr), value :: ctx, db
            integer(c_int) :: rc
        end function

        function kiwi_set_local_interpolation(ctx, bilinear) bind(C, name="kiwi_set_local_interpolation") result(rc)
            import :: c_ptr, c_int
            type(c_ptr), value :: ctx
            integer(c_int), value :: bilinear
            integer(c_int) :: rc
        end function

        function kiwi_set_spacial_undersampling(ctx, xunder, zunder) bind(C, name="kiwi_set_spacial_undersampling") result(rc)
            import :: c_ptr, c_int
            type(c_ptr), value :: ctx
            integer(c_int), value :: xunder, zunder
            integer(c_int) :: rc
        end function

        ! components: array of C strings, one per receiver (c_loc of null-terminated character buffers)
        function kiwi_set_receivers(ctx, n, lat_deg, lon_deg, depth, components) bind(C, name="kiwi_set_receivers") result(rc)
            import :: c_ptr, c_int, c_double, c_float
            type(c_ptr), value :: ctx
            integer(c_int), value :: n
            real(c_double), intent(in) :: lat_deg(*), lon_deg(*)
            real(c_float), intent(in) :: depth(*)
            type(c_ptr), intent(in) :: components(*)
            integer(c_int) :: rc
        end function

        function kiwi_switch_receiver(ctx, ireceiver, state) bind(C, name="kiwi_switch_receiver") result(rc)
            import :: c_ptr, c_int
            type(c_ptr), value :: ctx
            integer(c_int), value :: ireceiver, state
            integer(c_int) :: rc
        end function

        function kiwi_set_source_location(ctx, lat_deg, lon_deg, ref_time) bind(C, name="kiwi_set_source_location") result(rc)
            import :: c_ptr, c_int, c_float, c_double
            type(c_ptr), value :: ctx
            real(c_float), value :: lat_deg, lon_deg
            real(c_double), value :: ref_time
            integer(c_int) :: rc
        end function

        function kiwi_set_effective_dt(ctx, effective_dt) bind(C, name="kiwi_set_effective_dt") result(rc)
            import :: c_ptr, c_int, c_float
            type(c_ptr), value :: ctx
            real(c_float), value :: effective_dt
            integer(c_int) :: rc
        end function

        function kiwi_set_ref_seismogram(ctx, ireceiver, icomponent, tbegin, n, data) bind(C, name="kiwi_set_ref_seismogram") result(rc)
            import :: c_ptr, c_int, c_float
            type(c_ptr), value :: ctx
            integer(c_int), value :: ireceiver, icomponent, n
            real(c_float), value :: tbegin
            real(c_float), intent(in) :: data(*)
            integer(c_int) :: rc
        end function

        function kiwi_set_misfit_method(ctx, norm_id) bind(C, name="kiwi_set_misfit_method") result(rc)
            import :: c_ptr, c_int
            type(c_ptr), value :: ctx
            integer(c_int), value :: norm_id
            integer(c_int) :: rc
        end function

        function kiwi_set_misfit_taper(ctx, ireceiver, n, x, y) bind(C, name="kiwi_set_misfit_taper") result(rc)
            import :: c_ptr, c_int, c_float
            type(c_ptr), value :: ctx
            integer(c_int), value :: ireceiver, n
            real(c_float), intent(in) :: x(*), y(*)
            integer(c_int) :: rc
        end function

        function kiwi_set_misfit_filter(ctx, ireceiver, n, x, y) bind(C, name="kiwi_set_misfit_filter") result(rc)
            import :: c_ptr, c_int, c_float
            type(c_ptr), value :: ctx
            integer(c_int), value :: ireceiver, n
            real(c_float), intent(in) :: x(*), y(*)
            integer(c_int) :: rc
        end function

        function kiwi_set_synthetics_factor(ctx, factor) bind(C, name="kiwi_set_synthetics_factor") result(rc)
            import :: c_ptr, c_int, c_float
            type(c_ptr), value :: ctx
            real(c_float), value :: factor
            integer(c_int) :: rc
        end function

        function kiwi_set_floating_shiftrange(ctx, ireceiver, shift_lo, shift_hi) bind(C, name="kiwi_set_floating_shiftrange") result(rc)
            import :: c_ptr, c_int, c_float
            type(c_ptr), value :: ctx
            integer(c_int), value :: ireceiver
            real(c_float), value :: shift_lo, shift_hi
            integer(c_int) :: rc
        end function

        function kiwi_shift_ref_seismogram(ctx, ireceiver, shift) bind(C, name="kiwi_shift_ref_seismogram") result(rc)
            import :: c_ptr, c_int, c_float
            type(c_ptr), value :: ctx
            integer(c_int), value :: ireceiver
            real(c_float), value :: shift
            integer(c_int) :: rc
        end function

        function kiwi_autoshift_ref_seismogram(ctx, ireceiver, shift_lo, shift_hi, shifts, cap, n) &
                bind(C, name="kiwi_autoshift_ref_seismogram") result(rc)
            import :: c_ptr, c_int, c_float
            type(c_ptr), value :: ctx
            integer(c_int), value :: ireceiver, cap
            real(c_float), value :: shift_lo, shift_hi
            real(c_float), dimension(*), intent(out) :: shifts
            integer(c_int), intent(out) :: n
            integer(c_int) :: rc
        end function

        function kiwi_get_cross_correlations(ctx, ireceiver, shift_lo, shift_hi, cc, cap, ncomp, nshift) &
                bind(C, name="kiwi_get_cross_correlations") result(rc)
            import :: c_ptr, c_int, c_float
            type(c_ptr), value :: ctx
            integer(c_int), value :: ireceiver, cap
            real(c_float), value :: shift_lo, shift_hi
            real(c_float), dimension(*), intent(out) :: cc      ! (nshift, ncomp)
            integer(c_int), intent(out) :: ncomp, nshift
            integer(c_int) :: rc
        end function

        function kiwi_gfdb_interpolate(db, nipx, nipz, device) bind(C, name="kiwi_gfdb_interpolate") result(newdb)
            import :: c_ptr, c_int
            type(c_ptr), value :: db
            integer(c_int), value :: nipx, nipz, device
            type(c_ptr) :: newdb
        end function

        function kiwi_get_nmisfits(ctx) bind(C, name="kiwi_get_nmisfits") result(n)
            import :: c_ptr, c_int
            type(c_ptr), value :: ctx
            integer(c_int) :: n
        end function

        ! params(nparams, ns), misfits(2, nmisfits, ns), status(ns): Fortran column-major = C row-major transposed
        function kiwi_eval_sources(ctx, sourcetype, ns, nparams, params, misfits, status) bind(C, name="kiwi_eval_sources") result(rc)
            import :: c_ptr, c_int, c_float
            type(c_ptr), value :: ctx
            integer(c_int), value :: sourcetype, ns, nparams
            real(c_float), intent(in) :: params(*)
            real(c_float), intent(out) :: misfits(*)
            integer(c_int), intent(out) :: status(*)
            integer(c_int) :: rc
        end function

        function kiwi_set_source_params(ctx, sourcetype, nparams, params) bind(C, name="kiwi_set_source_params") result(rc)
            import :: c_ptr, c_int, c_float
            type(c_ptr), value :: ctx
            integer(c_int), value :: sourcetype, nparams
            real(c_float), intent(in) :: params(*)
            integer(c_int) :: rc
        end function

        function kiwi_get_misfits(ctx, misfits, cap_pairs, nmisfits) bind(C, name="kiwi_get_misfits") result(rc)
            import :: c_ptr, c_int, c_float
            type(c_ptr), value :: ctx
            real(c_float), intent(out) :: misfits(*)
            integer(c_int), value :: cap_pairs
            integer(c_int), intent(out) :: nmisfits
            integer(c_int) :: rc
        end function

        function kiwi_get_global_misfit(ctx, misfit) bind(C, name="kiwi_get_global_misfit") result(rc)
            import :: c_ptr, c_int, c_float
            type(c_ptr), value :: ctx
            real(c_float), intent(out) :: misfit
            integer(c_int) :: rc
        end function

        function kiwi_set_source_params_mask(ctx, mask, n) bind(C, name="kiwi_set_source_params_mask") result(rc)
            import :: c_ptr, c_int
            type(c_ptr), value :: ctx
            integer(c_int), intent(in) :: mask(*)          ! 0 / 1
            integer(c_int), value :: n
            integer(c_int) :: rc
        end function

        function kiwi_set_source_subparams(ctx, subparams, n) bind(C, name="kiwi_set_source_subparams") result(rc)
            import :: c_ptr, c_int, c_float
            type(c_ptr), value :: ctx
            real(c_float), intent(in) :: subparams(*)
            integer(c_int), value :: n
            integer(c_int) :: rc
        end function

        function kiwi_set_source_subparams_limits(ctx, mins, maxs, n) bind(C, name="kiwi_set_source_subparams_limits") result(rc)
            import :: c_ptr, c_int, c_float
            type(c_ptr), value :: ctx
            real(c_float), intent(in) :: mins(*), maxs(*)
            integer(c_int), value :: n
            integer(c_int) :: rc
        end function

        function kiwi_get_source_subparams(ctx, subparams, cap, n) bind(C, name="kiwi_get_source_subparams") result(rc)
            import :: c_ptr, c_int, c_float
            type(c_ptr), value :: ctx
            real(c_float), intent(out) :: subparams(*)
            integer(c_int), value :: cap
            integer(c_int), intent(out) :: n
            integer(c_int) :: rc
        end function

        ! replaces minimize_lm (minimizer_engine.f90:729-806): lmdif with the Jacobian sources as one batch
        function kiwi_minimize_lm(ctx, info, iterations, misfit) bind(C, name="kiwi_minimize_lm") result(rc)
            import :: c_ptr, c_int, c_float
            type(c_ptr), value :: ctx
            integer(c_int), intent(out) :: info, iterations
            real(c_float), intent(out) :: misfit
            integer(c_int) :: rc
        end function

        function kiwi_get_seismogram(ctx, ireceiver, icomponent, which, first_index, n, buf, cap) bind(C, name="kiwi_get_seismogram") result(rc)
            import :: c_ptr, c_int, c_float
            type(c_ptr), value :: ctx
            integer(c_int), value :: ireceiver, icomponent, which, cap
            integer(c_int), intent(out) :: first_index, n
            real(c_float), intent(out) :: buf(*)
            integer(c_int) :: rc
        end function

        ! ---- the remaining engine-facing entry points ------------------------------------------------------------------

        function kiwi_gfdb_read(path) bind(C, name="kiwi_gfdb_read") result(db)          ! KGF1 file of this library
            import :: c_ptr, c_char
            character(kind=c_char), intent(in) :: path(*)                                 ! NUL terminated
            type(c_ptr) :: db
        end function

        function kiwi_gfdb_read_hdf(basepath) bind(C, name="kiwi_gfdb_read_hdf") result(db)   ! <base>.index + <base>.<i>.chunk
            import :: c_ptr, c_char
            character(kind=c_char), intent(in) :: basepath(*)
            type(c_ptr) :: db
        end function

        subroutine kiwi_gfdb_destroy(db) bind(C, name="kiwi_gfdb_destroy")
            import :: c_ptr
            type(c_ptr), value :: db
        end subroutine

        function kiwi_set_crust2x2(ctx, path) bind(C, name="kiwi_set_crust2x2") result(rc)    ! crust2x2_load, minimizer.f90:1669-1674
            import :: c_ptr, c_int, c_char
            type(c_ptr), value :: ctx
            character(kind=c_char), intent(in) :: path(*)
            integer(c_int) :: rc
        end function

        function kiwi_set_source_constraints(ctx, n, points, normals) bind(C, name="kiwi_set_source_constraints") result(rc)
            import :: c_ptr, c_int, c_float
            type(c_ptr), value :: ctx
            integer(c_int), value :: n
            real(c_float), intent(in) :: points(3,*), normals(3,*)
            integer(c_int) :: rc
        end function

        function kiwi_set_source_crustal_thickness_limit(ctx, limit) bind(C, name="kiwi_set_source_crustal_thickness_limit") result(rc)
            import :: c_ptr, c_int, c_float
            type(c_ptr), value :: ctx
            real(c_float), value :: limit
            integer(c_int) :: rc
        end function

        function kiwi_get_source_crustal_thickness(ctx, thickness) bind(C, name="kiwi_get_source_crustal_thickness") result(rc)
            import :: c_ptr, c_int, c_float
            type(c_ptr), value :: ctx
            real(c_float), intent(out) :: thickness
            integer(c_int) :: rc
        end function

        function kiwi_get_distances(ctx, distances, azimuths, cap, n) bind(C, name="kiwi_get_distances") result(rc)
            import :: c_ptr, c_int, c_double
            type(c_ptr), value :: ctx
            real(c_double), intent(out) :: distances(*), azimuths(*)
            integer(c_int), value :: cap
            integer(c_int), intent(out) :: n
            integer(c_int) :: rc
        end function

        function kiwi_get_principal_axes(ctx, pax, tax) bind(C, name="kiwi_get_principal_axes") result(rc)
            import :: c_ptr, c_int, c_float
            type(c_ptr), value :: ctx
            real(c_float), intent(out) :: pax(2), tax(2)
            integer(c_int) :: rc
        end function

        function kiwi_get_floating_shifts(ctx, shifts, cap, n) bind(C, name="kiwi_get_floating_shifts") result(rc)   ! in samples
            import :: c_ptr, c_int
            type(c_ptr), value :: ctx
            integer(c_int), intent(out) :: shifts(*)
            integer(c_int), value :: cap
            integer(c_int), intent(out) :: n
            integer(c_int) :: rc
        end function

        function kiwi_get_peak_amplitudes(ctx, differentiate, maxabs, cap, n) bind(C, name="kiwi_get_peak_amplitudes") result(rc)
            import :: c_ptr, c_int, c_float
            type(c_ptr), value :: ctx
            integer(c_int), value :: differentiate, cap
            real(c_float), intent(out) :: maxabs(*)
            integer(c_int), intent(out) :: n
            integer(c_int) :: rc
        end function

        function kiwi_get_arias_intensities(ctx, intensities, cap, n) bind(C, name="kiwi_get_arias_intensities") result(rc)
            import :: c_ptr, c_int, c_float
            type(c_ptr), value :: ctx
            real(c_float), intent(out) :: intensities(*)
            integer(c_int), value :: cap
            integer(c_int), intent(out) :: n
            integer(c_int) :: rc
        end function

        ! probe_get / probe_get_amp_spectrum for output_seismograms and output_seismogram_spectra:
        ! which_probe 0 synthetics, 1 references; which_processing 0 plain, 1 tapered, 2 filtered
        function kiwi_get_probe(ctx, ireceiver, icomponent, which_probe, which_processing, first_index, n, buf, cap) &
                bind(C, name="kiwi_get_probe") result(rc)
            import :: c_ptr, c_int, c_float
            type(c_ptr), value :: ctx
            integer(c_int), value :: ireceiver, icomponent, which_probe, which_processing, cap
            integer(c_int), intent(out) :: first_index, n
            real(c_float), intent(out) :: buf(*)
            integer(c_int) :: rc
        end function

        function kiwi_get_probe_spectrum(ctx, ireceiver, icomponent, which_probe, which_processing, df, n, buf, cap) &
                bind(C, name="kiwi_get_probe_spectrum") result(rc)
            import :: c_ptr, c_int, c_float
            type(c_ptr), value :: ctx
            integer(c_int), value :: ireceiver, icomponent, which_probe, which_processing, cap
            real(c_float), intent(out) :: df
            integer(c_int), intent(out) :: n
            real(c_float), intent(out) :: buf(*)
            integer(c_int) :: rc
        end function

        ! batched ground-motion diagnostics: out(3, nenabled, ns) = peak velocity, peak acceleration, Arias intensity
        function kiwi_eval_ground_motion(ctx, sourcetype, ns, nparams, params, out, status) bind(C, name="kiwi_eval_ground_motion") result(rc)
            import :: c_ptr, c_int, c_float
            type(c_ptr), value :: ctx
            integer(c_int), value :: sourcetype, ns, nparams
            real(c_float), intent(in) :: params(*)
            real(c_float), intent(out) :: out(*)
            integer(c_int), intent(out) :: status(*)
            integer(c_int) :: rc
        end function

        ! make_global_misfits + best candidate on the device (seismosizer.py:843-922); pass c_null_ptr for d_misfits to use the block
        ! the last kiwi_eval_sources call left on the GPU, and for the optional arrays that are not wanted
        function kiwi_outer_misfits(ctx, ns, d_misfits, receiver_weights, outer_norm, anarchy, nboot, bweights, misfits_by_s, best, best_value) &
                bind(C, name="kiwi_outer_misfits") result(rc)
            import :: c_ptr, c_int
            type(c_ptr), value :: ctx, d_misfits, receiver_weights, bweights, misfits_by_s, best, best_value
            integer(c_int), value :: ns, outer_norm, anarchy, nboot
            integer(c_int) :: rc
        end function

        function kiwi_set_mt_grid(ctx, enabled) bind(C, name="kiwi_set_mt_grid") result(rc)
            import :: c_ptr, c_int
            type(c_ptr), value :: ctx
            integer(c_int), value :: enabled
            integer(c_int) :: rc
        end function

        function kiwi_set_accumulation(ctx, reference_order) bind(C, name="kiwi_set_accumulation") result(rc)
            import :: c_ptr, c_int
            type(c_ptr), value :: ctx
            integer(c_int), value :: reference_order
            integer(c_int) :: rc
        end function

        function kiwi_set_eikonal_device(ctx, min_batch) bind(C, name="kiwi_set_eikonal_device") result(rc)
            import :: c_ptr, c_int
            type(c_ptr), value :: ctx
            integer(c_int), value :: min_batch
            integer(c_int) :: rc
        end function

        function kiwi_set_share_syntheses(ctx, enabled) bind(C, name="kiwi_set_share_syntheses") result(rc)
            import :: c_ptr, c_int
            type(c_ptr), value :: ctx
            integer(c_int), value :: enabled
            integer(c_int) :: rc
        end function


    end interface

end module
