!-----------------------------------------------------------------------------------------------
! sw_interface_cuda.f90 -- thin iso_c_binding shim that plugs libswcuda.so (include/swcuda.h) into
! the reference's algorithm layer in place of gpu/interface/sw_interface_gpu.f90 +
! gpu/kernel/*.f90.  SOURCE ONLY: this image has no Fortran compiler, so the file is not built or
! tested here (the same C ABI is exercised through ctypes in tests/).  Nothing in the reference's
! algorithm layer (control/, core/) changes except the three call sites listed in INTEGRATION.md.
!
!   swcuda_c_binding                      bind(C) interfaces of the WHOLE C ABI (separate, generated file)
!   shallow_water_interface_cuda_module   init_device_data_cuda, expl_shallow_water_cuda,
!                                         download_ssh_cuda, finalize_device_data_cuda, and the
!                                         14 envoke_<name>_kernel_cuda / _sync_cuda pairs
!
! Across MPI ranks the cut is y-slabs (1 x nranks process grid), one block per rank; the halo
! exchange then happens inside swcu_step over NCCL.  A single rank may instead own any
! bppnx x bppny blocks (dealt over the visible GPUs): they are linked pairwise and step together
! through swcu_step_group, which pulls halos device-to-device.
!-----------------------------------------------------------------------------------------------
! (module swcuda_c_binding: fortran/swcuda_c_binding.f90, generated from include/swcuda.h by
! fortran/gen_bindings.py -- every entry point of the C ABI, the constants and the structs)


module shallow_water_interface_cuda_module
    use iso_c_binding
    use swcuda_c_binding
    use kind_module, only: wp8 => SHR_KIND_R8, wp4 => SHR_KIND_R4
    use mpp_module
    use decomposition_module, only: domain_type, domain => domain_data
    use ocean_module, only: ocean_type, ocean_data
    use grid_module, only: grid_type, grid_data
    use config_sw_module, only: full_free_surface, time_smooth, trans_terms, ksw_lat, use_tracers, tracer_num
    use errors_module, only: abort_model
    use kernel_interface_module, only: kernel_parameters_type
    use mpp_sync_module, only: sync_parameters_type
    implicit none
    save
    private

    type(c_ptr), allocatable :: ctx(:)      ! one resident context per local block

    public :: init_device_data_cuda, expl_shallow_water_cuda, download_ssh_cuda, finalize_device_data_cuda
    public :: envoke_sw_update_ssh_kernel_cuda, envoke_sw_update_ssh_sync_cuda
    public :: envoke_hh_update_kernel_cuda, envoke_hh_update_sync_cuda
    public :: envoke_uv_trans_vort_kernel_cuda, envoke_uv_trans_vort_sync_cuda
    public :: envoke_uv_trans_kernel_cuda, envoke_uv_trans_sync_cuda
    public :: envoke_stress_components_kernel_cuda, envoke_stress_components_sync_cuda
    public :: envoke_uv_diff2_kernel_cuda, envoke_uv_diff2_sync_cuda
    public :: envoke_sw_update_uv_kernel_cuda, envoke_sw_update_uv_sync_cuda
    public :: envoke_sw_next_step_kernel_cuda, envoke_sw_next_step_sync_cuda
    public :: envoke_hh_shift_kernel_cuda, envoke_hh_shift_sync_cuda
    public :: envoke_hh_init_kernel_cuda, envoke_hh_init_sync_cuda
    public :: envoke_check_ssh_err_kernel_cuda, envoke_check_ssh_err_sync_cuda
    public :: envoke_tran_diff_fluxes_kernel_cuda, envoke_tran_diff_fluxes_sync_cuda
    public :: envoke_tran_diff_tracer_kernel_cuda, envoke_tran_diff_tracer_sync_cuda
    public :: envoke_tracer_next_step_kernel_cuda, envoke_tracer_next_step_sync_cuda

contains

    ! text of swcu_last_error() as a Fortran string
    function last_error() result(msg)
        character(len=:), allocatable :: msg
        character(kind=c_char), pointer :: p(:)
        type(c_ptr) :: cp
        integer :: n
        cp = swcu_last_error()
        msg = ''
        if (.not. c_associated(cp)) return
        call c_f_pointer(cp, p, [1024])
        n = 0
        do while (n < 1024)
            if (p(n + 1) == c_null_char) exit
            n = n + 1
        enddo
        allocate(character(len=n) :: msg)
        msg = transfer(p(1:n), msg)
    end function

    subroutine check(rc, what)
        integer(c_int), intent(in) :: rc
        character(*), intent(in) :: what
        if (rc /= SWCU_OK) call abort_model('swcuda: '//what//': '//last_error())      ! shared/errors.f90:30-37
    end subroutine

    subroutine up8(k, field, a)
        integer, intent(in) :: k
        integer(c_int), intent(in) :: field
        real(wp8), target, intent(in) :: a(:, :)
        call check(swcu_upload(ctx(k), field, c_loc(a)), 'upload')
    end subroutine

    subroutine up4(k, field, a)
        integer, intent(in) :: k
        integer(c_int), intent(in) :: field
        real(wp4), target, intent(in) :: a(:, :)
        call check(swcu_upload(ctx(k), field, c_loc(a)), 'upload')
    end subroutine

    ! .true. if local blocks k1 and k2 are side or corner neighbours in the block grid
    logical function blocks_touch(k1, k2)
        integer, intent(in) :: k1, k2
        integer :: dx, dy
        dx = rel(domain%bnx_start(k1), domain%bnx_end(k1), domain%bnx_start(k2), domain%bnx_end(k2))
        dy = rel(domain%bny_start(k1), domain%bny_end(k1), domain%bny_start(k2), domain%bny_end(k2))
        blocks_touch = dx /= 2 .and. dy /= 2 .and. (dx /= 0 .or. dy /= 0)
    contains
        integer function rel(a0, a1, b0, b1)   ! +1 / -1 adjacent, 0 same range, 2 unrelated
            integer, intent(in) :: a0, a1, b0, b1
            rel = 2
            if (b0 == a1 + 1) rel = 1
            if (b1 + 1 == a0) rel = -1
            if (b0 == a0 .and. b1 == a1) rel = 0
        end function
    end function

    ! replaces init_device_data (control/init_data.f90:127-145): called once after init_grid_data /
    ! init_ocean_data have filled the host arrays
    subroutine init_device_data_cuda()
        type(swcu_dims) :: d
        type(swcu_params) :: p
        character(kind=c_char), target :: id(128)
        character(kind=c_char), target :: blob_mine(SWCU_PEER_BLOB_BYTES), blob_nbr(SWCU_PEER_BLOB_BYTES)
        integer :: k, k2, t, ierr, ndev, node_comm, node_size, node_rank, side, nbr, to, dev

        ndev = max(1, int(swcu_device_count()))
        if (mpp_count > 1 .and. domain%bcount > 1) call abort_model('swcuda: several ranks need one block per rank')
        ! which GPU: ranks of one node take the node's GPUs in the order of their NODE-LOCAL rank (one block per
        ! rank); a single rank deals its blocks over the visible GPUs
        node_rank = 0; node_size = 1; node_comm = mpi_comm_null
        if (mpp_count > 1) then
            call mpi_comm_split_type(mpp_cart_comm, mpi_comm_type_shared, 0, mpi_info_null, node_comm, ierr)
            call mpi_comm_size(node_comm, node_size, ierr)
            call mpi_comm_rank(node_comm, node_rank, ierr)
            if (node_size > ndev) call abort_model('swcuda: more ranks on this node than GPUs')
        endif
        allocate(ctx(domain%bcount))
        p%full_free_surface = full_free_surface; p%trans_terms = trans_terms; p%ksw_lat = ksw_lat
        p%time_smooth = time_smooth; p%use_tracers = use_tracers; p%mode = SWCU_MODE_FUSED
        do k = 1, domain%bcount
            d%nx_start = domain%bnx_start(k); d%nx_end = domain%bnx_end(k)
            d%ny_start = domain%bny_start(k); d%ny_end = domain%bny_end(k)
            d%bnd_x1 = domain%bbnd_x1(k); d%bnd_x2 = domain%bbnd_x2(k)
            d%bnd_y1 = domain%bbnd_y1(k); d%bnd_y2 = domain%bbnd_y2(k)
            dev = mod(k - 1, ndev)
            if (mpp_count > 1) dev = mod(node_rank, ndev)
            call check(swcu_create(ctx(k), d, p, int(dev, c_int)), 'create')
            call up4(k, SWCU_F_LU,  grid_data%lu %block(k)%field); call up4(k, SWCU_F_LUU, grid_data%luu%block(k)%field)
            call up4(k, SWCU_F_LUH, grid_data%luh%block(k)%field); call up4(k, SWCU_F_LCU, grid_data%lcu%block(k)%field)
            call up4(k, SWCU_F_LCV, grid_data%lcv%block(k)%field); call up4(k, SWCU_F_LLU, grid_data%llu%block(k)%field)
            call up4(k, SWCU_F_LLV, grid_data%llv%block(k)%field)
            call up4(k, SWCU_F_DX,  grid_data%dx %block(k)%field); call up4(k, SWCU_F_DY,  grid_data%dy %block(k)%field)
            call up4(k, SWCU_F_DXT, grid_data%dxt%block(k)%field); call up4(k, SWCU_F_DYT, grid_data%dyt%block(k)%field)
            call up4(k, SWCU_F_DXH, grid_data%dxh%block(k)%field); call up4(k, SWCU_F_DYH, grid_data%dyh%block(k)%field)
            call up4(k, SWCU_F_DXB, grid_data%dxb%block(k)%field); call up4(k, SWCU_F_DYB, grid_data%dyb%block(k)%field)
            call up4(k, SWCU_F_RLH_S, grid_data%rlh_s%block(k)%field)
            call up8(k, SWCU_F_HHQ_REST, grid_data%hhq_rest%block(k)%field)
            call up8(k, SWCU_F_SSH,    ocean_data%ssh   %block(k)%field); call up8(k, SWCU_F_SSHP,   ocean_data%sshp  %block(k)%field)
            call up8(k, SWCU_F_UBRTR,  ocean_data%ubrtr %block(k)%field); call up8(k, SWCU_F_UBRTRP, ocean_data%ubrtrp%block(k)%field)
            call up8(k, SWCU_F_VBRTR,  ocean_data%vbrtr %block(k)%field); call up8(k, SWCU_F_VBRTRP, ocean_data%vbrtrp%block(k)%field)
            call up8(k, SWCU_F_MU,     ocean_data%mu    %block(k)%field)
            if (use_tracers > 0) then   ! expl_tracer runs inside the step (control/tracer.f90:44-61), over all tracers
                if (tracer_num > 1) call check(swcu_set_option(ctx(k), 'tracer_num'//c_null_char, int(tracer_num, c_int)), 'tracer_num')
                do t = 1, tracer_num
                    call check(swcu_set_option(ctx(k), 'tracer_select'//c_null_char, int(t - 1, c_int)), 'tracer_select')
                    call up8(k, SWCU_F_FF1,  ocean_data%ff1(t) %block(k)%field)
                    call up8(k, SWCU_F_FF1P, ocean_data%ff1p(t)%block(k)%field)
                enddo
                call check(swcu_set_option(ctx(k), 'tracer_select'//c_null_char, 0_c_int), 'tracer_select')
            endif
            call check(swcu_envoke_hh_init(ctx(k)), 'hh_init')
        enddo
        ! several blocks in this rank: link every pair of side / corner neighbours once
        do k = 1, domain%bcount
            do k2 = k + 1, domain%bcount
                if (blocks_touch(k, k2)) call check(swcu_link(ctx(k), ctx(k2)), 'link')
            enddo
        enddo
        ! The fused step reads the inputs TWO layers outside a block's interior; the reference's arrays are valid
        ! one layer out (include/swcuda.h, swcu_widen_halos).  Linked blocks fetch the second layer now, ranks
        ! below, over the communicator.
        do k = 1, domain%bcount
            call check(swcu_widen_halos(ctx(k)), 'widen_halos')
        enddo
        if (mpp_count > 1) then
            ! one NCCL communicator over the y-slab ranks; the id travels over MPI
            if (mpp_is_master()) call check(swcu_comm_unique_id(c_loc(id)), 'unique_id')
            call mpi_bcast(id, 128, mpi_character, 0, mpp_cart_comm, ierr)
            call check(swcu_comm_init(ctx(1), int(mpp_count, c_int), int(mpp_rank, c_int), c_loc(id)), 'comm_init')
            call check(swcu_widen_halos(ctx(1)), 'widen_halos')
            if (node_size == mpp_count .and. (use_tracers == 0 .or. tracer_num == 1)) then
                ! all ranks on one node (the peer path carries one tracer): from here on halo rows are stored straight into the neighbours' memory
                ! (CUDA IPC, fused into the step kernel); the 2048-byte blobs travel once over MPI
                call check(swcu_comm_destroy(ctx(1)), 'comm_destroy')
                call check(swcu_peer_export(ctx(1), c_loc(blob_mine)), 'peer_export')
                do side = 0, 1
                    nbr = mpp_rank - 1 + 2 * side                       ! side 0: rank-1, side 1: rank+1
                    to  = mpp_rank + 1 - 2 * side
                    if (nbr < 0 .or. nbr >= mpp_count) nbr = mpi_proc_null
                    if (to  < 0 .or. to  >= mpp_count) to  = mpi_proc_null
                    call mpi_sendrecv(blob_mine, SWCU_PEER_BLOB_BYTES, mpi_character, to,  side,   &
                                      blob_nbr,  SWCU_PEER_BLOB_BYTES, mpi_character, nbr, side,   &
                                      mpp_cart_comm, mpi_status_ignore, ierr)
                    if (nbr /= mpi_proc_null) call check(swcu_peer_attach(ctx(1), int(side, c_int), c_loc(blob_nbr)), 'peer_attach')
                enddo
            endif
        endif
    end subroutine

    ! replaces expl_shallow_water_gpu (control/shallow_water/shallow_water.f90:184-251); model.f90:146
    ! calls it from inside the long-lived !$omp parallel region, so only the master thread launches
    subroutine expl_shallow_water_cuda(tau)
        real(wp8), intent(in) :: tau
        !$omp master
        if (domain%bcount > 1) then
            call check(swcu_step_group(ctx, int(domain%bcount, c_int), real(tau, c_double), 1_c_int), 'step_group')
        else
            call check(swcu_step(ctx(1), real(tau, c_double), 1_c_int), 'step')
        endif
        !$omp end master
        !$omp barrier
    end subroutine

    ! replaces ocean_data%ssh%sync_host_device(domain, .false.) before local_output (model.f90:181);
    ! also polls the device-side check_ssh_err flag (K11)
    subroutine download_ssh_cuda()
        integer :: k
        integer(c_long) :: bad
        do k = 1, domain%bcount
            call check(swcu_synchronize(ctx(k), bad), 'synchronize (SWCU_ERR_BLOWUP = the reference''s SIGFPRE predict error, vel_ssh.f90:57)')
            call check(swcu_download(ctx(k), SWCU_F_SSH, c_loc(ocean_data%ssh%block(k)%field)), 'download')
        enddo
    end subroutine

    !------------------------------------------------------------------------------------------
    ! Per-kernel route (SWCU_MODE_REFERENCE): procedures with the abstract interfaces
    ! envoke_empty_kernel / envoke_empty_sync (core/kernel_interface.f90:38-46), so the reference's
    ! expl_shallow_water / expl_tracer can keep their envoke(sub_kernel, sub_sync, parameters) calls
    ! and only point sub_kernel / sub_sync at these instead of the CPU binders
    ! (control/shallow_water/shallow_water.f90:36-92).  k = -1 in a sync means "all blocks"
    ! (core/kernel_interface.f90:101).
    !------------------------------------------------------------------------------------------
    subroutine sync_all(kernel_id, k)
        integer(c_int), intent(in) :: kernel_id
        integer, intent(in) :: k
        integer :: kk
        if (k >= 1) then
            call check(swcu_envoke_sync(ctx(k), kernel_id), 'sync')
        else
            !$omp master
            do kk = 1, domain%bcount
                call check(swcu_envoke_sync(ctx(kk), kernel_id), 'sync')
            enddo
            !$omp end master
            !$omp barrier
        endif
    end subroutine

    subroutine envoke_sw_update_ssh_kernel_cuda(k, param)
        integer, intent(in) :: k
        type(kernel_parameters_type), intent(in) :: param
        call check(swcu_envoke_kernel(ctx(k), SWCU_K_SW_UPDATE_SSH, real(param%tau, c_double)), 'sw_update_ssh')
    end subroutine
    subroutine envoke_sw_update_ssh_sync_cuda(k, sync_parameters)
        integer, intent(in) :: k
        type(sync_parameters_type), intent(in) :: sync_parameters
        call sync_all(SWCU_K_SW_UPDATE_SSH, k)
    end subroutine

    subroutine envoke_hh_update_kernel_cuda(k, param)
        integer, intent(in) :: k
        type(kernel_parameters_type), intent(in) :: param
        call check(swcu_envoke_kernel(ctx(k), SWCU_K_HH_UPDATE, real(param%tau, c_double)), 'hh_update')
    end subroutine
    subroutine envoke_hh_update_sync_cuda(k, sync_parameters)
        integer, intent(in) :: k
        type(sync_parameters_type), intent(in) :: sync_parameters
        call sync_all(SWCU_K_HH_UPDATE, k)
    end subroutine

    subroutine envoke_uv_trans_vort_kernel_cuda(k, param)
        integer, intent(in) :: k
        type(kernel_parameters_type), intent(in) :: param
        call check(swcu_envoke_kernel(ctx(k), SWCU_K_UV_TRANS_VORT, real(param%tau, c_double)), 'uv_trans_vort')
    end subroutine
    subroutine envoke_uv_trans_vort_sync_cuda(k, sync_parameters)
        integer, intent(in) :: k
        type(sync_parameters_type), intent(in) :: sync_parameters
        call sync_all(SWCU_K_UV_TRANS_VORT, k)
    end subroutine

    subroutine envoke_uv_trans_kernel_cuda(k, param)
        integer, intent(in) :: k
        type(kernel_parameters_type), intent(in) :: param
        call check(swcu_envoke_kernel(ctx(k), SWCU_K_UV_TRANS, real(param%tau, c_double)), 'uv_trans')
    end subroutine
    subroutine envoke_uv_trans_sync_cuda(k, sync_parameters)
        integer, intent(in) :: k
        type(sync_parameters_type), intent(in) :: sync_parameters
        call sync_all(SWCU_K_UV_TRANS, k)
    end subroutine

    subroutine envoke_stress_components_kernel_cuda(k, param)
        integer, intent(in) :: k
        type(kernel_parameters_type), intent(in) :: param
        call check(swcu_envoke_kernel(ctx(k), SWCU_K_STRESS_COMPONENTS, real(param%tau, c_double)), 'stress_components')
    end subroutine
    subroutine envoke_stress_components_sync_cuda(k, sync_parameters)
        integer, intent(in) :: k
        type(sync_parameters_type), intent(in) :: sync_parameters
        call sync_all(SWCU_K_STRESS_COMPONENTS, k)
    end subroutine

    subroutine envoke_uv_diff2_kernel_cuda(k, param)
        integer, intent(in) :: k
        type(kernel_parameters_type), intent(in) :: param
        call check(swcu_envoke_kernel(ctx(k), SWCU_K_UV_DIFF2, real(param%tau, c_double)), 'uv_diff2')
    end subroutine
    subroutine envoke_uv_diff2_sync_cuda(k, sync_parameters)
        integer, intent(in) :: k
        type(sync_parameters_type), intent(in) :: sync_parameters
        call sync_all(SWCU_K_UV_DIFF2, k)
    end subroutine

    subroutine envoke_sw_update_uv_kernel_cuda(k, param)
        integer, intent(in) :: k
        type(kernel_parameters_type), intent(in) :: param
        call check(swcu_envoke_kernel(ctx(k), SWCU_K_SW_UPDATE_UV, real(param%tau, c_double)), 'sw_update_uv')
    end subroutine
    subroutine envoke_sw_update_uv_sync_cuda(k, sync_parameters)
        integer, intent(in) :: k
        type(sync_parameters_type), intent(in) :: sync_parameters
        call sync_all(SWCU_K_SW_UPDATE_UV, k)
    end subroutine

    subroutine envoke_sw_next_step_kernel_cuda(k, param)
        integer, intent(in) :: k
        type(kernel_parameters_type), intent(in) :: param
        call check(swcu_envoke_kernel(ctx(k), SWCU_K_SW_NEXT_STEP, real(param%tau, c_double)), 'sw_next_step')
    end subroutine
    subroutine envoke_sw_next_step_sync_cuda(k, sync_parameters)
        integer, intent(in) :: k
        type(sync_parameters_type), intent(in) :: sync_parameters
        call sync_all(SWCU_K_SW_NEXT_STEP, k)
    end subroutine

    subroutine envoke_hh_shift_kernel_cuda(k, param)
        integer, intent(in) :: k
        type(kernel_parameters_type), intent(in) :: param
        call check(swcu_envoke_kernel(ctx(k), SWCU_K_HH_SHIFT, real(param%tau, c_double)), 'hh_shift')
    end subroutine
    subroutine envoke_hh_shift_sync_cuda(k, sync_parameters)
        integer, intent(in) :: k
        type(sync_parameters_type), intent(in) :: sync_parameters
        call sync_all(SWCU_K_HH_SHIFT, k)
    end subroutine

    subroutine envoke_hh_init_kernel_cuda(k, param)
        integer, intent(in) :: k
        type(kernel_parameters_type), intent(in) :: param
        call check(swcu_envoke_kernel(ctx(k), SWCU_K_HH_INIT, real(param%tau, c_double)), 'hh_init')
    end subroutine
    subroutine envoke_hh_init_sync_cuda(k, sync_parameters)
        integer, intent(in) :: k
        type(sync_parameters_type), intent(in) :: sync_parameters
        call sync_all(SWCU_K_HH_INIT, k)
    end subroutine

    subroutine envoke_check_ssh_err_kernel_cuda(k, param)
        integer, intent(in) :: k
        type(kernel_parameters_type), intent(in) :: param
        call check(swcu_envoke_kernel(ctx(k), SWCU_K_CHECK_SSH_ERR, real(param%tau, c_double)), 'check_ssh_err')
    end subroutine
    subroutine envoke_check_ssh_err_sync_cuda(k, sync_parameters)
        integer, intent(in) :: k
        type(sync_parameters_type), intent(in) :: sync_parameters
        call sync_all(SWCU_K_CHECK_SSH_ERR, k)
    end subroutine

    subroutine envoke_tran_diff_fluxes_kernel_cuda(k, param)
        integer, intent(in) :: k
        type(kernel_parameters_type), intent(in) :: param
        call check(swcu_envoke_kernel(ctx(k), SWCU_K_TRAN_DIFF_FLUXES, real(param%tau, c_double)), 'tran_diff_fluxes')
    end subroutine
    subroutine envoke_tran_diff_fluxes_sync_cuda(k, sync_parameters)
        integer, intent(in) :: k
        type(sync_parameters_type), intent(in) :: sync_parameters
        call sync_all(SWCU_K_TRAN_DIFF_FLUXES, k)
    end subroutine

    subroutine envoke_tran_diff_tracer_kernel_cuda(k, param)
        integer, intent(in) :: k
        type(kernel_parameters_type), intent(in) :: param
        call check(swcu_envoke_kernel(ctx(k), SWCU_K_TRAN_DIFF_TRACER, real(param%tau, c_double)), 'tran_diff_tracer')
    end subroutine
    subroutine envoke_tran_diff_tracer_sync_cuda(k, sync_parameters)
        integer, intent(in) :: k
        type(sync_parameters_type), intent(in) :: sync_parameters
        call sync_all(SWCU_K_TRAN_DIFF_TRACER, k)
    end subroutine

    subroutine envoke_tracer_next_step_kernel_cuda(k, param)
        integer, intent(in) :: k
        type(kernel_parameters_type), intent(in) :: param
        call check(swcu_envoke_kernel(ctx(k), SWCU_K_TRACER_NEXT_STEP, real(param%tau, c_double)), 'tracer_next_step')
    end subroutine
    subroutine envoke_tracer_next_step_sync_cuda(k, sync_parameters)
        integer, intent(in) :: k
        type(sync_parameters_type), intent(in) :: sync_parameters
        call sync_all(SWCU_K_TRACER_NEXT_STEP, k)
    end subroutine

    subroutine finalize_device_data_cuda()
        integer :: k
        do k = 1, size(ctx)
            call check(swcu_destroy(ctx(k)), 'destroy')
        enddo
        deallocate(ctx)
    end subroutine

end module shallow_water_interface_cuda_module
