!  obsgrid_gpu_iface.F90 -- the Fortran side of the drop-in boundary.
!
!  OBSGRID's driver, namelist handling, little_r ingest and netCDF output stay in Fortran.
!  The three procedures its stage drivers call for the hot path keep their names, argument
!  order and meaning (so proc_oa.F90:728-736 and proc_qc.F90:372-391 do not change) and
!  forward to libobsgrid_gpu.so through ISO_C_BINDING:
!
!     cressman     src/oa_module.F90:25-32     ->  og_cressman
!     error_max    src/qc1_module.F90:9-15     ->  og_error_max
!     buddy_check  src/qc2_module.F90:9-14     ->  og_buddy_check
!
!  Build: compile this file instead of the bodies of oa_module.F90 (cressman), qc1_module.F90
!  and qc2_module.F90 (see INTEGRATION.md), link with -lobsgrid_gpu.  The image this library
!  is developed in has no Fortran compiler, so this file is reviewed, not compiled, there; the
!  C ABI it binds (include/obsgrid_gpu.h) is exercised by the ctypes tests with the very same
!  argument lists.
!
!  Conventions (include/obsgrid_gpu.h): arrays are passed as they are (column-major
!  gridded(iew,jns), 1-based REAL(4) coordinates), LOGICALs as INTEGER(C_INT) 0/1,
!  CHARACTER(8) variable names as integer codes, every C function returns 0 or a negative
!  status and never STOPs -- the wrappers turn a failure into the reference's fatal
!  error_handler call (src/error_handler.F90:53-92).

MODULE obsgrid_gpu

   USE , INTRINSIC :: ISO_C_BINDING
   IMPLICIT NONE
   PRIVATE
   PUBLIC :: og_start , og_stop , og_handle , og_var_code

   TYPE ( C_PTR ) , SAVE :: og_handle = C_NULL_PTR

   !  The per-time batched API (include/obsgrid_gpu.h, INTEGRATION.md section 2): interoperable
   !  mirrors of og_qc_params (namelist record 4), og_oa_params (records 7-9) and og_time_desc.
   !  Member order and types are the C structs'; BIND(C) gives them the C compiler's padding
   !  (tests/test_abi.py compares the member lists).
   INTEGER , PARAMETER :: OG_MAX_SCAN = 10            ! src/proc_oa.F90:105

   TYPE , BIND ( C ) :: og_qc_params
      INTEGER ( C_INT ) :: qc_test_error_max , qc_test_buddy
      REAL ( C_FLOAT ) :: max_error_t , max_error_uv , max_error_rh , max_error_dewpoint , max_error_p
      REAL ( C_FLOAT ) :: max_buddy_t , max_buddy_uv , max_buddy_rh , max_buddy_dewpoint , max_buddy_p
      REAL ( C_FLOAT ) :: buddy_weight
      INTEGER ( C_INT ) :: time_hhmmss
      INTEGER ( C_INT ) :: qc_dewpoint
      INTEGER ( C_INT ) :: var_mask
   END TYPE og_qc_params

   TYPE , BIND ( C ) :: og_oa_params
      INTEGER ( C_INT ) :: radius_influence ( OG_MAX_SCAN )
      REAL ( C_FLOAT ) :: radius_influence_sfc_mult
      INTEGER ( C_INT ) :: use_first_guess
      INTEGER ( C_INT ) :: scale_cressman_rh_decreases
      INTEGER ( C_INT ) :: smooth_type
      INTEGER ( C_INT ) :: smooth_sfc_wind , smooth_sfc_temp , smooth_sfc_rh , smooth_sfc_slp
      INTEGER ( C_INT ) :: smooth_upper_wind , smooth_upper_temp , smooth_upper_rh
      INTEGER ( C_INT ) :: surface_only
      INTEGER ( C_INT ) :: clamp_fg_rh
      INTEGER ( C_INT ) :: var_mask
   END TYPE og_oa_params

   !  pointers: C_LOC of the caller's arrays -- t(jns,iew,kbu) etc. as src/first_guess.inc declares
   !  them, the observation table rows as INTEGRATION.md section 2 describes
   TYPE , BIND ( C ) :: og_time_desc
      INTEGER ( C_INT ) :: iew , jns , kbu
      REAL ( C_FLOAT ) :: dx_km
      TYPE ( C_PTR ) :: pressure
      TYPE ( C_PTR ) :: t , u , v , rh
      TYPE ( C_PTR ) :: slp
      INTEGER ( C_INT ) :: nrep
      TYPE ( C_PTR ) :: xob , yob
      TYPE ( C_PTR ) :: lon , elev
      TYPE ( C_PTR ) :: ob_u , ob_v , ob_t , ob_rh
      TYPE ( C_PTR ) :: ob_slp
      TYPE ( C_PTR ) :: qc_u , qc_v , qc_t , qc_rh
      TYPE ( C_PTR ) :: qc_slp
      TYPE ( C_PTR ) :: ob_td        ! optional rows (C_NULL_PTR: absent), include/obsgrid_gpu.h
      TYPE ( C_PTR ) :: qc_td
      TYPE ( C_PTR ) :: bogus
   END TYPE og_time_desc

   PUBLIC :: OG_MAX_SCAN , og_qc_params , og_oa_params , og_time_desc

   INTERFACE

      INTEGER ( C_INT ) FUNCTION og_init ( ctx , device ) BIND ( C , NAME = 'og_init' )
         IMPORT :: C_PTR , C_INT
         TYPE ( C_PTR ) , INTENT ( OUT ) :: ctx
         INTEGER ( C_INT ) , VALUE :: device
      END FUNCTION og_init

      INTEGER ( C_INT ) FUNCTION og_finalize ( ctx ) BIND ( C , NAME = 'og_finalize' )
         IMPORT :: C_PTR , C_INT
         TYPE ( C_PTR ) , VALUE :: ctx
      END FUNCTION og_finalize

      FUNCTION og_last_error ( ) BIND ( C , NAME = 'og_last_error' ) RESULT ( msg )
         IMPORT :: C_PTR
         TYPE ( C_PTR ) :: msg
      END FUNCTION og_last_error

      INTEGER ( C_INT ) FUNCTION og_cressman ( ctx , diff , xob , yob , n , gridded , iew , jns , &
      crsdot , var , radius_influence , dxd_km , u_banana , v_banana , pressure_hpa , passes , &
      smooth_type , use_first_guess , fg_value , scale_cressman_rh_decreases ) &
      BIND ( C , NAME = 'og_cressman' )
         IMPORT :: C_PTR , C_INT , C_FLOAT
         TYPE ( C_PTR ) , VALUE :: ctx
         REAL ( C_FLOAT ) , INTENT ( IN ) :: diff ( * ) , xob ( * ) , yob ( * )
         INTEGER ( C_INT ) , VALUE :: n
         REAL ( C_FLOAT ) , INTENT ( INOUT ) :: gridded ( * )
         INTEGER ( C_INT ) , VALUE :: iew , jns , crsdot , var , radius_influence
         REAL ( C_FLOAT ) , VALUE :: dxd_km
         REAL ( C_FLOAT ) , INTENT ( IN ) :: u_banana ( * ) , v_banana ( * )
         REAL ( C_FLOAT ) , VALUE :: pressure_hpa
         INTEGER ( C_INT ) , VALUE :: passes , smooth_type , use_first_guess
         REAL ( C_FLOAT ) , INTENT ( IN ) :: fg_value ( * )
         INTEGER ( C_INT ) , VALUE :: scale_cressman_rh_decreases
      END FUNCTION og_cressman

      INTEGER ( C_INT ) FUNCTION og_error_max ( ctx , obs , xob , yob , station_elevation , &
      qc_flag , n , gridded , iew , jns , var , max_difference , pressure_hpa , local_time , &
      aob , err , thr ) BIND ( C , NAME = 'og_error_max' )
         IMPORT :: C_PTR , C_INT , C_FLOAT
         TYPE ( C_PTR ) , VALUE :: ctx
         REAL ( C_FLOAT ) , INTENT ( IN ) :: obs ( * ) , xob ( * ) , yob ( * ) , station_elevation ( * )
         INTEGER ( C_INT ) , INTENT ( INOUT ) :: qc_flag ( * )
         INTEGER ( C_INT ) , VALUE :: n
         REAL ( C_FLOAT ) , INTENT ( IN ) :: gridded ( * )
         INTEGER ( C_INT ) , VALUE :: iew , jns , var
         REAL ( C_FLOAT ) , VALUE :: max_difference , pressure_hpa
         INTEGER ( C_INT ) , INTENT ( IN ) :: local_time ( * )
         TYPE ( C_PTR ) , VALUE :: aob , err , thr          ! nullable diagnostics
      END FUNCTION og_error_max

      INTEGER ( C_INT ) FUNCTION og_buddy_check ( ctx , obs , n , xob , yob , qc_flag , &
      station_elevation , gridded , iew , jns , var , pressure_hpa , dx_km , buddy_weight , &
      buddy_difference , aob , err , difmax , buddy_num_final ) BIND ( C , NAME = 'og_buddy_check' )
         IMPORT :: C_PTR , C_INT , C_FLOAT
         TYPE ( C_PTR ) , VALUE :: ctx
         REAL ( C_FLOAT ) , INTENT ( IN ) :: obs ( * )
         INTEGER ( C_INT ) , VALUE :: n
         REAL ( C_FLOAT ) , INTENT ( IN ) :: xob ( * ) , yob ( * )
         INTEGER ( C_INT ) , INTENT ( INOUT ) :: qc_flag ( * )
         REAL ( C_FLOAT ) , INTENT ( IN ) :: station_elevation ( * ) , gridded ( * )
         INTEGER ( C_INT ) , VALUE :: iew , jns , var
         REAL ( C_FLOAT ) , VALUE :: pressure_hpa , dx_km , buddy_weight , buddy_difference
         TYPE ( C_PTR ) , VALUE :: aob , err , difmax , buddy_num_final   ! nullable diagnostics
      END FUNCTION og_buddy_check

      INTEGER ( C_INT ) FUNCTION og_ob_density ( ctx , xob , yob , grid_dist , numobs , tobbox , &
      iew , jns ) BIND ( C , NAME = 'og_ob_density' )
         IMPORT :: C_PTR , C_INT , C_FLOAT
         TYPE ( C_PTR ) , VALUE :: ctx
         REAL ( C_FLOAT ) , INTENT ( IN ) :: xob ( * ) , yob ( * )
         REAL ( C_FLOAT ) , VALUE :: grid_dist
         INTEGER ( C_INT ) , VALUE :: numobs
         REAL ( C_FLOAT ) , INTENT ( INOUT ) :: tobbox ( * )
         INTEGER ( C_INT ) , VALUE :: iew , jns
      END FUNCTION og_ob_density

      INTEGER ( C_INT ) FUNCTION og_obs_distance ( ctx , tobbox , distance , iew , jns , dx ) &
      BIND ( C , NAME = 'og_obs_distance' )
         IMPORT :: C_PTR , C_INT , C_FLOAT
         TYPE ( C_PTR ) , VALUE :: ctx
         REAL ( C_FLOAT ) , INTENT ( IN ) :: tobbox ( * )
         REAL ( C_FLOAT ) , INTENT ( OUT ) :: distance ( * )
         INTEGER ( C_INT ) , VALUE :: iew , jns
         REAL ( C_FLOAT ) , VALUE :: dx
      END FUNCTION og_obs_distance

      INTEGER ( C_INT ) FUNCTION og_wind_increment_to_c_grid ( ctx , u , uA , uC , v , vA , vC , &
      jns , iew , kbu ) BIND ( C , NAME = 'og_wind_increment_to_c_grid' )
         IMPORT :: C_PTR , C_INT , C_FLOAT
         TYPE ( C_PTR ) , VALUE :: ctx
         REAL ( C_FLOAT ) , INTENT ( INOUT ) :: u ( * ) , v ( * )
         REAL ( C_FLOAT ) , INTENT ( IN ) :: uA ( * ) , uC ( * ) , vA ( * ) , vC ( * )
         INTEGER ( C_INT ) , VALUE :: jns , iew , kbu
      END FUNCTION og_wind_increment_to_c_grid

      INTEGER ( C_INT ) FUNCTION og_pres_sfc_from_slp ( ctx , pres_sfc , slp_C , slp_x , out , &
      jns , iew ) BIND ( C , NAME = 'og_pres_sfc_from_slp' )
         IMPORT :: C_PTR , C_INT , C_FLOAT
         TYPE ( C_PTR ) , VALUE :: ctx
         REAL ( C_FLOAT ) , INTENT ( IN ) :: pres_sfc ( * ) , slp_C ( * ) , slp_x ( * )
         REAL ( C_FLOAT ) , INTENT ( OUT ) :: out ( * )
         INTEGER ( C_INT ) , VALUE :: jns , iew
      END FUNCTION og_pres_sfc_from_slp

      INTEGER ( C_INT ) FUNCTION og_temporal_interp ( ctx , first , second , out , jns , iew , nlev , &
      icount_fdda , icount_1 , icount_2 , cross ) BIND ( C , NAME = 'og_temporal_interp' )
         IMPORT :: C_PTR , C_INT , C_FLOAT
         TYPE ( C_PTR ) , VALUE :: ctx
         REAL ( C_FLOAT ) , INTENT ( IN ) :: first ( * ) , second ( * )
         REAL ( C_FLOAT ) , INTENT ( OUT ) :: out ( * )
         INTEGER ( C_INT ) , VALUE :: jns , iew , nlev , icount_fdda , icount_1 , icount_2 , cross
      END FUNCTION og_temporal_interp

      INTEGER ( C_INT ) FUNCTION og_mixing_ratio ( ctx , rh , t , h , terrain , slp_x , pressure , &
      iew , jns , kbu , qv , psfc ) BIND ( C , NAME = 'og_mixing_ratio' )
         IMPORT :: C_PTR , C_INT , C_FLOAT
         TYPE ( C_PTR ) , VALUE :: ctx
         REAL ( C_FLOAT ) , INTENT ( INOUT ) :: rh ( * ) , qv ( * ) , psfc ( * )
         REAL ( C_FLOAT ) , INTENT ( IN ) :: t ( * ) , h ( * ) , terrain ( * ) , slp_x ( * ) , pressure ( * )
         INTEGER ( C_INT ) , VALUE :: iew , jns , kbu
      END FUNCTION og_mixing_ratio

      !  proc_qc / proc_oa of one analysis time over the flat observation table
      INTEGER ( C_INT ) FUNCTION og_qc_time ( ctx , d , p ) BIND ( C , NAME = 'og_qc_time' )
         IMPORT :: C_PTR , C_INT , og_time_desc , og_qc_params
         TYPE ( C_PTR ) , VALUE :: ctx
         TYPE ( og_time_desc ) , INTENT ( IN ) :: d
         TYPE ( og_qc_params ) , INTENT ( IN ) :: p
      END FUNCTION og_qc_time

      INTEGER ( C_INT ) FUNCTION og_qc_time_fdda ( ctx , d , p , tobbox , odis ) BIND ( C , NAME = 'og_qc_time_fdda' )
         IMPORT :: C_PTR , C_INT , C_FLOAT , og_time_desc , og_qc_params
         TYPE ( C_PTR ) , VALUE :: ctx
         TYPE ( og_time_desc ) , INTENT ( IN ) :: d
         TYPE ( og_qc_params ) , INTENT ( IN ) :: p
         REAL ( C_FLOAT ) , INTENT ( INOUT ) :: tobbox ( * )
         REAL ( C_FLOAT ) , INTENT ( OUT ) :: odis ( * )
      END FUNCTION og_qc_time_fdda

      INTEGER ( C_INT ) FUNCTION og_oa_time ( ctx , d , p , k_begin , k_end ) BIND ( C , NAME = 'og_oa_time' )
         IMPORT :: C_PTR , C_INT , og_time_desc , og_oa_params
         TYPE ( C_PTR ) , VALUE :: ctx
         TYPE ( og_time_desc ) , INTENT ( IN ) :: d
         TYPE ( og_oa_params ) , INTENT ( IN ) :: p
         INTEGER ( C_INT ) , VALUE :: k_begin , k_end
      END FUNCTION og_oa_time

      INTEGER ( C_INT ) FUNCTION og_qc_oa_time ( ctx , d , qp , op , k_begin , k_end , pipeline_groups ) &
      BIND ( C , NAME = 'og_qc_oa_time' )
         IMPORT :: C_PTR , C_INT , og_time_desc , og_qc_params , og_oa_params
         TYPE ( C_PTR ) , VALUE :: ctx
         TYPE ( og_time_desc ) , INTENT ( IN ) :: d
         TYPE ( og_qc_params ) , INTENT ( IN ) :: qp
         TYPE ( og_oa_params ) , INTENT ( IN ) :: op
         INTEGER ( C_INT ) , VALUE :: k_begin , k_end , pipeline_groups
      END FUNCTION og_qc_oa_time

      !  0 = OG_ARITH_EXACT (default), 1 = OG_ARITH_CONTRACTED
      INTEGER ( C_INT ) FUNCTION og_set_arithmetic ( ctx , mode ) BIND ( C , NAME = 'og_set_arithmetic' )
         IMPORT :: C_PTR , C_INT
         TYPE ( C_PTR ) , VALUE :: ctx
         INTEGER ( C_INT ) , VALUE :: mode
      END FUNCTION og_set_arithmetic

      !  One slab over several GPUs (include/obsgrid_gpu.h): installs the all-reduce callback
      !     INTEGER(C_INT) FUNCTION fn ( user , buf , count , dtype , op , stream ) BIND ( C )
      !  (all arguments by VALUE: C_PTR , C_PTR , C_SIZE_T , C_INT , C_INT , C_PTR); pass it as
      !  C_FUNLOC ( fn ).  An MPI host wraps ncclAllReduce there (INTEGRATION.md, section 5), with
      !  rank / world = its MPI rank and size.  world <= 1 or C_NULL_FUNPTR: back to one GPU.
      INTEGER ( C_INT ) FUNCTION og_set_exchange ( ctx , rank , world , fn , user ) BIND ( C , NAME = 'og_set_exchange' )
         IMPORT :: C_PTR , C_INT , C_FUNPTR
         TYPE ( C_PTR ) , VALUE :: ctx
         INTEGER ( C_INT ) , VALUE :: rank , world
         TYPE ( C_FUNPTR ) , VALUE :: fn
         TYPE ( C_PTR ) , VALUE :: user
      END FUNCTION og_set_exchange

   END INTERFACE

   PUBLIC :: og_cressman , og_error_max , og_buddy_check , og_ob_density , og_obs_distance , og_fail
   PUBLIC :: og_wind_increment_to_c_grid , og_pres_sfc_from_slp , og_mixing_ratio , og_temporal_interp
   PUBLIC :: og_qc_time , og_qc_time_fdda , og_oa_time , og_qc_oa_time , og_set_arithmetic , og_set_exchange

CONTAINS

   !  One context per process, created before the first time period (driver.F90, before the
   !  time loop) and destroyed after the last.
   SUBROUTINE og_start ( device )
      INTEGER , INTENT ( IN ) :: device
      IF ( og_init ( og_handle , INT ( device , C_INT ) ) .NE. 0 ) CALL og_fail ( 'og_init' )
   END SUBROUTINE og_start

   SUBROUTINE og_stop
      INTEGER ( C_INT ) :: rc
      rc = og_finalize ( og_handle )
      og_handle = C_NULL_PTR
   END SUBROUTINE og_stop

   !  CHARACTER(8) variable name -> og_var code (include/obsgrid_gpu.h: OG_VAR_*).
   INTEGER ( C_INT ) FUNCTION og_var_code ( name )
      CHARACTER ( LEN = 8 ) , INTENT ( IN ) :: name
      SELECT CASE ( name )
         CASE ( 'UU      ' ) ; og_var_code = 1
         CASE ( 'VV      ' ) ; og_var_code = 2
         CASE ( 'TT      ' ) ; og_var_code = 3
         CASE ( 'RH      ' ) ; og_var_code = 4
         CASE ( 'PMSL    ' ) ; og_var_code = 5
         CASE ( 'PSFC    ' , 'PMSLPSFC' ) ; og_var_code = 6
         CASE ( 'DEWPOINT' ) ; og_var_code = 7
         CASE DEFAULT ; og_var_code = -1
      END SELECT
   END FUNCTION og_var_code

   !  The library never STOPs; the reference's convention for an unrecoverable error is a
   !  fatal error_handler call (src/error_handler.F90:53-92).
   SUBROUTINE og_fail ( where )
      CHARACTER ( LEN = * ) , INTENT ( IN ) :: where
      INCLUDE 'error.inc'
      INTERFACE
         INCLUDE 'error.int'
      END INTERFACE
      CHARACTER ( KIND = C_CHAR ) , POINTER :: cmsg ( : )
      CHARACTER ( LEN = 200 ) :: msg
      INTEGER :: i
      msg = ' '
      CALL C_F_POINTER ( og_last_error ( ) , cmsg , (/ 200 /) )
      DO i = 1 , 200
         IF ( cmsg ( i ) .EQ. C_NULL_CHAR ) EXIT
         msg ( i : i ) = cmsg ( i )
      END DO
      error_number = 00363001
      error_message ( 1 : 31 ) = 'obsgrid_gpu                    '
      error_message ( 32 : ) = ' ' // where // ': ' // TRIM ( msg )
      fatal = .TRUE.
      listing = .FALSE.
      CALL error_handler ( error_number , error_message , fatal , listing )
   END SUBROUTINE og_fail

END MODULE obsgrid_gpu

!  ---------------------------------------------------------------------------------------
!  Replacement bodies with the reference's exact interfaces.  They live in the modules the
!  callers already USE (oa_module, qc1, qc2): see INTEGRATION.md for the three-line patch that
!  swaps each body for an INCLUDE of the matching block below.
!  ---------------------------------------------------------------------------------------

!  MODULE oa_module :: cressman          (replaces src/oa_module.F90:25-218)
SUBROUTINE cressman ( obs_value , obs , xob , yob , numobs , &
gridded , iew , jns , &
crsdot , name , radius_influence , &
dxd , u_banana , v_banana , pressure , passes , smooth_type , lat_center , &
use_first_guess , fg_value , scale_cressman_rh_decreases )

   USE , INTRINSIC :: ISO_C_BINDING
   USE obsgrid_gpu
   IMPLICIT NONE

   INTEGER , INTENT ( IN )                             :: numobs
   INTEGER , INTENT ( IN )                             :: iew , jns
   REAL , INTENT ( IN ) , DIMENSION ( numobs )         :: obs_value , obs , xob , yob
   REAL , INTENT ( IN ) , DIMENSION ( numobs )         :: fg_value
   REAL , INTENT ( INOUT ) , DIMENSION ( iew , jns )   :: gridded
   INTEGER , INTENT ( IN )                             :: radius_influence
   CHARACTER ( 8 ) , INTENT ( IN )                     :: name
   REAL , INTENT ( IN ) , DIMENSION ( iew , jns )      :: u_banana , v_banana
   REAL , INTENT ( IN )                                :: dxd , pressure
   INTEGER                                             :: crsdot
   INTEGER , INTENT ( IN )                             :: passes , smooth_type
   REAL , INTENT ( IN )                                :: lat_center
   LOGICAL , INTENT ( IN )                             :: use_first_guess
   LOGICAL                                             :: scale_cressman_rh_decreases

   INTEGER ( C_INT ) :: rc

   !  obs_value and lat_center are unused by the reference body as well.
   rc = og_cressman ( og_handle , obs , xob , yob , INT ( numobs , C_INT ) , gridded , &
                      INT ( iew , C_INT ) , INT ( jns , C_INT ) , INT ( crsdot , C_INT ) , &
                      og_var_code ( name ) , INT ( radius_influence , C_INT ) , dxd , &
                      u_banana , v_banana , pressure , INT ( passes , C_INT ) , &
                      INT ( smooth_type , C_INT ) , MERGE ( 1_C_INT , 0_C_INT , use_first_guess ) , &
                      fg_value , MERGE ( 1_C_INT , 0_C_INT , scale_cressman_rh_decreases ) )
   IF ( rc .NE. 0 ) CALL og_fail ( 'cressman' )

END SUBROUTINE cressman

!  MODULE qc1 :: error_max               (replaces src/qc1_module.F90:9-300)
SUBROUTINE error_max ( station_id , obs , ship_report , xob , yob ,  &
station_elevation , &
index , qc_flag , numobs , &
gridded ,  iew , jns ,  &
name , max_difference , pressure , local_time , print_error_max )

   USE , INTRINSIC :: ISO_C_BINDING
   USE obsgrid_gpu
   IMPLICIT NONE

   INTEGER                                        :: numobs
   CHARACTER ( LEN =   8 ) , DIMENSION ( : )      :: station_id
   REAL                    , DIMENSION ( : )      :: obs , xob , yob
   REAL                    , DIMENSION ( : )      :: station_elevation
   LOGICAL                 , DIMENSION ( : )      :: ship_report
   INTEGER                 , DIMENSION ( : )      :: index
   INTEGER                 , DIMENSION ( : )      :: qc_flag
   REAL                    , DIMENSION( : , : )   :: gridded
   INTEGER                                        :: iew , jns
   CHARACTER ( LEN =   8 )                        :: name
   INTEGER                 , DIMENSION ( : )      :: local_time
   REAL                                           :: max_difference
   REAL                                           :: pressure
   LOGICAL                                        :: print_error_max

   REAL ( C_FLOAT ) , DIMENSION ( numobs ) , TARGET :: aob , err , thr
   INTEGER ( C_INT ) :: rc
   INTEGER :: num

   !  Assumed-shape dummies may be non-contiguous: hand contiguous sections to C (the compiler
   !  copies in/out when it has to; proc_qc passes whole automatic arrays, so it does not).
   rc = og_error_max ( og_handle , obs ( 1 : numobs ) , xob ( 1 : numobs ) , yob ( 1 : numobs ) , &
                       station_elevation ( 1 : numobs ) , qc_flag ( 1 : numobs ) , &
                       INT ( numobs , C_INT ) , gridded ( 1 : iew , 1 : jns ) , &
                       INT ( iew , C_INT ) , INT ( jns , C_INT ) , og_var_code ( name ) , &
                       max_difference , pressure , local_time ( 1 : numobs ) , &
                       C_LOC ( aob ) , C_LOC ( err ) , C_LOC ( thr ) )
   IF ( rc .NE. 0 ) CALL og_fail ( 'error_max' )

   !  The reference's optional listing (src/qc1_module.F90:251-296) from the returned
   !  diagnostics; station_id / ship_report / index only ever fed this print.
   IF ( print_error_max ) THEN
      DO num = 1 , numobs
         IF ( ABS ( err ( num ) ) .GT. thr ( num ) ) THEN
            WRITE ( UNIT = * , FMT = '(A8,1X,A8,1X,I5,4(1X,F12.3))' ) name , station_id ( num ) , &
               index ( num ) , obs ( num ) , aob ( num ) , err ( num ) , thr ( num )
         END IF
      END DO
   END IF

END SUBROUTINE error_max

!  MODULE qc2 :: buddy_check             (replaces src/qc2_module.F90:9-388)
SUBROUTINE buddy_check ( station_id , obs , numobs , xob , yob , qc_flag ,   &
station_elevation , &
gridded , iew , jns , name ,                      &
pressure , local_time , dx , buddy_weight , buddy_difference , print_buddy )

   USE , INTRINSIC :: ISO_C_BINDING
   USE obsgrid_gpu
   IMPLICIT NONE

   INTEGER                                        :: numobs
   CHARACTER ( LEN =   8 ) , DIMENSION ( : )      :: station_id
   REAL                    , DIMENSION ( : )      :: obs , xob , yob
   INTEGER                 , DIMENSION ( : )      :: qc_flag
   REAL                    , DIMENSION ( : )      :: station_elevation
   REAL                    , DIMENSION( : , : )   :: gridded
   INTEGER                                        :: iew , jns
   CHARACTER ( LEN =   8 )                        :: name
   REAL                                           :: buddy_weight ,   &
                                                     buddy_difference , &
                                                     dx          , &
                                                     pressure
   INTEGER                 , DIMENSION ( : )      :: local_time
   LOGICAL                                        :: print_buddy

   REAL ( C_FLOAT ) , DIMENSION ( numobs ) , TARGET    :: aob , err , difmax
   INTEGER ( C_INT ) , DIMENSION ( numobs ) , TARGET   :: buddy_num_final
   INTEGER ( C_INT ) :: rc
   INTEGER :: num

   !  local_time is not used by the reference body either (src/qc2_module.F90:9-388).
   rc = og_buddy_check ( og_handle , obs ( 1 : numobs ) , INT ( numobs , C_INT ) , &
                         xob ( 1 : numobs ) , yob ( 1 : numobs ) , qc_flag ( 1 : numobs ) , &
                         station_elevation ( 1 : numobs ) , gridded ( 1 : iew , 1 : jns ) , &
                         INT ( iew , C_INT ) , INT ( jns , C_INT ) , og_var_code ( name ) , &
                         pressure , dx , buddy_weight , buddy_difference , &
                         C_LOC ( aob ) , C_LOC ( err ) , C_LOC ( difmax ) , C_LOC ( buddy_num_final ) )
   IF ( rc .NE. 0 ) CALL og_fail ( 'buddy_check' )

   IF ( print_buddy ) THEN   ! src/qc2_module.F90:339-384
      DO num = 1 , numobs
         IF ( ABS ( err ( num ) ) .GT. difmax ( num ) ) THEN
            WRITE ( UNIT = * , FMT = '(A8,1X,A8,1X,3(1X,F12.3),1X,I6)' ) name , station_id ( num ) , &
               obs ( num ) , err ( num ) , difmax ( num ) , buddy_num_final ( num )
         END IF
      END DO
   END IF

END SUBROUTINE buddy_check

!  MODULE qc0 :: ob_density / obs_distance   (replace src/qc0_module.F90:81-119 and :123-232)
SUBROUTINE ob_density ( xob , yob , grid_dist , numobs , tobbox , iew , jns )

   USE , INTRINSIC :: ISO_C_BINDING
   USE obsgrid_gpu
   IMPLICIT NONE

   REAL    , DIMENSION ( : )      :: xob , yob
   REAL                           :: grid_dist
   INTEGER                        :: numobs
   REAL    , DIMENSION ( : , : )  :: tobbox
   INTEGER                        :: iew , jns

   IF ( og_ob_density ( og_handle , xob ( 1 : numobs ) , yob ( 1 : numobs ) , grid_dist , &
                        INT ( numobs , C_INT ) , tobbox ( 1 : jns , 1 : iew ) , &
                        INT ( iew , C_INT ) , INT ( jns , C_INT ) ) .NE. 0 ) CALL og_fail ( 'ob_density' )

END SUBROUTINE ob_density

SUBROUTINE obs_distance ( tobbox , distance , iew , jns , dx )

   USE , INTRINSIC :: ISO_C_BINDING
   USE obsgrid_gpu
   IMPLICIT NONE

   INTEGER                                 :: jns , iew
   REAL , INTENT ( IN )                    :: dx
   REAL , DIMENSION ( : , : ) , INTENT ( IN )  :: tobbox
   REAL , DIMENSION ( : , : ) , INTENT ( OUT ) :: distance

   IF ( og_obs_distance ( og_handle , tobbox ( 1 : jns , 1 : iew ) , distance ( 1 : jns , 1 : iew ) , &
                          INT ( iew , C_INT ) , INT ( jns , C_INT ) , dx ) .NE. 0 ) CALL og_fail ( 'obs_distance' )

END SUBROUTINE obs_distance

!  MODULE final_analysis :: mixing_ratio   (replaces src/final_analysis_module.F90:1026-1201,
!  sfcprs included; called from proc_oa.F90:346)
SUBROUTINE mixing_ratio ( rh , t , h , terrain , slp_x , pressure , iew_alloc , jns_alloc , kbu_alloc , &
                          qv , psfc )

   USE , INTRINSIC :: ISO_C_BINDING
   USE obsgrid_gpu
   IMPLICIT NONE

   INTEGER , INTENT ( IN ) :: iew_alloc , jns_alloc , kbu_alloc
   REAL , INTENT ( IN ) , DIMENSION ( kbu_alloc ) :: pressure
   REAL , INTENT ( IN ) , DIMENSION ( jns_alloc , iew_alloc ) :: terrain , slp_x
   REAL , INTENT ( INOUT ) , DIMENSION ( jns_alloc , iew_alloc , kbu_alloc ) :: rh
   REAL , INTENT ( IN ) , DIMENSION ( jns_alloc , iew_alloc , kbu_alloc ) :: t , h
   REAL , INTENT ( INOUT ) , DIMENSION ( jns_alloc , iew_alloc , kbu_alloc ) :: qv
   REAL , INTENT ( INOUT ) , DIMENSION ( jns_alloc , iew_alloc ) :: psfc

   !  a missing 850/700/500 hPa level is the reference's STOP 'No_find_levels' (:1122-1127)
   IF ( og_mixing_ratio ( og_handle , rh , t , h , terrain , slp_x , pressure , INT ( iew_alloc , C_INT ) , &
                          INT ( jns_alloc , C_INT ) , INT ( kbu_alloc , C_INT ) , qv , psfc ) .NE. 0 ) &
      CALL og_fail ( 'mixing_ratio' )

END SUBROUTINE mixing_ratio

!  write_analysis, src/final_analysis_module.F90:316-319 and :336-339: the two statements
!     pertubation = u - uA ; CALL crs2dot_pertU ( pertubation , ... ) ; u = uC + pertubation
!  (and the V twin) become one call
SUBROUTINE wind_increment_to_c_grid ( u , uA , uC , v , vA , vC , jns_alloc , iew_alloc , kbu_alloc )

   USE , INTRINSIC :: ISO_C_BINDING
   USE obsgrid_gpu
   IMPLICIT NONE

   INTEGER , INTENT ( IN ) :: jns_alloc , iew_alloc , kbu_alloc
   REAL , INTENT ( INOUT ) , DIMENSION ( jns_alloc , iew_alloc , kbu_alloc ) :: u , v
   REAL , INTENT ( IN ) , DIMENSION ( jns_alloc , iew_alloc , kbu_alloc ) :: uA , uC , vA , vC

   IF ( og_wind_increment_to_c_grid ( og_handle , u , uA , uC , v , vA , vC , INT ( jns_alloc , C_INT ) , &
                                      INT ( iew_alloc , C_INT ) , INT ( kbu_alloc , C_INT ) ) .NE. 0 ) &
      CALL og_fail ( 'wind_increment_to_c_grid' )

END SUBROUTINE wind_increment_to_c_grid

!  write_analysis, src/final_analysis_module.F90:364-370 (inside IF ( .NOT. oa_psfc ))
SUBROUTINE pres_sfc_from_slp ( pres_sfc , slp_C , slp_x , out , jns_alloc , iew_alloc )

   USE , INTRINSIC :: ISO_C_BINDING
   USE obsgrid_gpu
   IMPLICIT NONE

   INTEGER , INTENT ( IN ) :: jns_alloc , iew_alloc
   REAL , INTENT ( IN ) , DIMENSION ( jns_alloc , iew_alloc ) :: pres_sfc , slp_C , slp_x
   REAL , INTENT ( OUT ) , DIMENSION ( jns_alloc , iew_alloc ) :: out

   IF ( og_pres_sfc_from_slp ( og_handle , pres_sfc , slp_C , slp_x , out , INT ( jns_alloc , C_INT ) , &
                               INT ( iew_alloc , C_INT ) ) .NE. 0 ) CALL og_fail ( 'pres_sfc_from_slp' )

END SUBROUTINE pres_sfc_from_slp

!  temporal_interp, src/tinterp_module.F90:126-213: each branch of the_3d_search / the_2d_search
!  becomes one call, e.g. for 'TT      ' (cross = 1; UU, VV, UA, VA, UC, VC pass cross = 0):
!     CALL temporal_interp_field ( all_3d(loop_count,first_time)%array , &
!                                  all_3d(loop_count,second_time)%array , t , jns_alloc , iew_alloc , &
!                                  kbu_alloc , icount_fdda , icount_1 , icount_2 , 1 )
SUBROUTINE temporal_interp_field ( first , second , out , jns_alloc , iew_alloc , nlev , &
                                   icount_fdda , icount_1 , icount_2 , cross )

   USE , INTRINSIC :: ISO_C_BINDING
   USE obsgrid_gpu
   IMPLICIT NONE

   INTEGER , INTENT ( IN ) :: jns_alloc , iew_alloc , nlev , icount_fdda , icount_1 , icount_2 , cross
   REAL , INTENT ( IN ) , DIMENSION ( jns_alloc , iew_alloc , nlev ) :: first , second
   REAL , INTENT ( OUT ) , DIMENSION ( jns_alloc , iew_alloc , nlev ) :: out

   IF ( og_temporal_interp ( og_handle , first , second , out , INT ( jns_alloc , C_INT ) , &
                             INT ( iew_alloc , C_INT ) , INT ( nlev , C_INT ) , INT ( icount_fdda , C_INT ) , &
                             INT ( icount_1 , C_INT ) , INT ( icount_2 , C_INT ) , INT ( cross , C_INT ) ) .NE. 0 ) &
      CALL og_fail ( 'temporal_interp' )

END SUBROUTINE temporal_interp_field

!------------------------------------------------------------------------------------------------
!  The batched path from the reference's own data structure (SURVEY 8f rank 1).
!
!  og_time_desc wants one flat observation table per analysis time; OBSGRID holds TYPE(report)
!  records with a linked list of TYPE(measurement) (src/obs_sort_module.F90:75-245) and derives
!  every slab's observations with proc_ob_access / query_ob / store_ob.  This module flattens the
!  list into the plain arrays of include/obsgrid_reports.h, lets og_reports_to_table decide -- once
!  per time -- what the per-slab calls decide again and again, and scatters flags and dew points back
!  (og_table_to_reports), leaving the list as proc_qc + qc_consistency + proc_oa leave it.
!
!     CALL og_flatten_reports ( obs , number_of_obs , reports , meas , nmeas )
!     ... og_reports_to_table ( reports , n , meas , nmeas , date , time , request_p_diff , desc , meas_index )
!     ... og_qc_time / og_oa_time / og_qc_oa_time ( desc ) ...
!     ... og_table_to_reports ( reports , n , meas , nmeas , date , time , desc , meas_index )
!     CALL og_unflatten_reports ( obs , number_of_obs , reports , meas )
!
!  replaces, in driver.F90, the calls of proc_qc (:455-459), qc_consistency and proc_oa's observation
!  access for one time period (INTEGRATION.md section 2 shows the patch).
MODULE obsgrid_gpu_reports

   USE , INTRINSIC :: ISO_C_BINDING
   USE observation                       !  TYPE(report), TYPE(measurement): src/obs_sort_module.F90
   USE map_utils                         !  latlon_to_ij
   USE map_utils_helper                  !  projd
   USE obsgrid_gpu , ONLY : og_time_desc
   IMPLICIT NONE
   PRIVATE
   PUBLIC :: og_field , og_meas , og_report , og_flatten_reports , og_unflatten_reports , &
             og_reports_to_table , og_table_to_reports

   !  include/obsgrid_reports.h, member for member
   TYPE , BIND ( C ) :: og_field
      REAL ( C_FLOAT ) :: data
      INTEGER ( C_INT ) :: qc
   END TYPE og_field

   TYPE , BIND ( C ) :: og_meas
      TYPE ( og_field ) :: pressure , height , temperature , dew_point , speed , direction , u , v , rh , thickness
   END TYPE og_meas

   TYPE , BIND ( C ) :: og_report
      REAL ( C_FLOAT ) :: latitude , longitude
      REAL ( C_FLOAT ) :: x , y
      REAL ( C_FLOAT ) :: elevation
      CHARACTER ( KIND = C_CHAR ) :: platform ( 40 )
      CHARACTER ( KIND = C_CHAR ) :: date_char ( 14 )
      INTEGER ( C_INT ) :: discard , bogus
      TYPE ( og_field ) :: slp
      TYPE ( og_field ) :: psfc
      INTEGER ( C_INT ) :: first_meas , n_meas
   END TYPE og_report

   INTERFACE
      INTEGER ( C_INT ) FUNCTION og_reports_to_table ( reports , nrep , meas , nmeas , date , time , &
      request_p_diff , table , meas_index ) BIND ( C , NAME = 'og_reports_to_table' )
         IMPORT :: C_INT , og_report , og_meas , og_time_desc
         TYPE ( og_report ) , INTENT ( INOUT ) :: reports ( * )
         INTEGER ( C_INT ) , VALUE :: nrep , nmeas , date , time , request_p_diff
         TYPE ( og_meas ) , INTENT ( INOUT ) :: meas ( * )
         TYPE ( og_time_desc ) , INTENT ( INOUT ) :: table
         INTEGER ( C_INT ) , INTENT ( OUT ) :: meas_index ( * )
      END FUNCTION og_reports_to_table

      INTEGER ( C_INT ) FUNCTION og_table_to_reports ( reports , nrep , meas , nmeas , date , time , &
      table , meas_index ) BIND ( C , NAME = 'og_table_to_reports' )
         IMPORT :: C_INT , og_report , og_meas , og_time_desc
         TYPE ( og_report ) , INTENT ( INOUT ) :: reports ( * )
         INTEGER ( C_INT ) , VALUE :: nrep , nmeas , date , time
         TYPE ( og_meas ) , INTENT ( INOUT ) :: meas ( * )
         TYPE ( og_time_desc ) , INTENT ( IN ) :: table
         INTEGER ( C_INT ) , INTENT ( IN ) :: meas_index ( * )
      END FUNCTION og_table_to_reports

      !  table entries whose measurement also answers a lower level (0: the table path is exact)
      INTEGER ( C_INT ) FUNCTION og_table_shared_levels ( meas_index , kbu , nrep ) BIND ( C , NAME = 'og_table_shared_levels' )
         IMPORT :: C_INT
         INTEGER ( C_INT ) , DIMENSION ( * ) , INTENT ( IN ) :: meas_index
         INTEGER ( C_INT ) , VALUE :: kbu , nrep
      END FUNCTION og_table_shared_levels
   END INTERFACE

CONTAINS

   PURE FUNCTION to_field ( f ) RESULT ( g )
      TYPE ( field ) , INTENT ( IN ) :: f
      TYPE ( og_field ) :: g
      g%data = f%data
      g%qc   = f%qc
   END FUNCTION to_field

   !  The report list as plain arrays, in list order.  meas is allocated here (one entry per
   !  measurement of every report); x, y are latlon_to_ij of the report's location, as
   !  proc_ob_access computes them per request (src/proc_ob_access.F90:204-208).
   SUBROUTINE og_flatten_reports ( obs , number_of_obs , reports , meas , nmeas )
      TYPE ( report ) , INTENT ( IN ) , DIMENSION ( : ) :: obs
      INTEGER , INTENT ( IN ) :: number_of_obs
      TYPE ( og_report ) , INTENT ( OUT ) , DIMENSION ( : ) :: reports
      TYPE ( og_meas ) , ALLOCATABLE , INTENT ( OUT ) , DIMENSION ( : ) :: meas
      INTEGER , INTENT ( OUT ) :: nmeas
      TYPE ( measurement ) , POINTER :: cur
      INTEGER :: i , c , m

      nmeas = 0
      DO i = 1 , number_of_obs
         cur => obs(i)%surface
         DO WHILE ( ASSOCIATED ( cur ) )
            nmeas = nmeas + 1
            cur => cur%next
         END DO
      END DO
      ALLOCATE ( meas ( MAX ( nmeas , 1 ) ) )

      m = 0
      DO i = 1 , number_of_obs
         reports(i)%latitude  = obs(i)%location%latitude
         reports(i)%longitude = obs(i)%location%longitude
         CALL latlon_to_ij ( projd , obs(i)%location%latitude , obs(i)%location%longitude , &
                             reports(i)%x , reports(i)%y )
         reports(i)%elevation = obs(i)%info%elevation
         DO c = 1 , 40
            reports(i)%platform(c) = obs(i)%info%platform(c:c)
         END DO
         DO c = 1 , 14
            reports(i)%date_char(c) = obs(i)%valid_time%date_char(c:c)
         END DO
         reports(i)%discard = MERGE ( 1 , 0 , obs(i)%info%discard )
         reports(i)%bogus   = MERGE ( 1 , 0 , obs(i)%info%bogus )
         reports(i)%slp  = to_field ( obs(i)%ground%slp )
         reports(i)%psfc = to_field ( obs(i)%ground%psfc )
         reports(i)%first_meas = m                  !  0-based index into meas
         reports(i)%n_meas = 0
         cur => obs(i)%surface
         DO WHILE ( ASSOCIATED ( cur ) )
            m = m + 1
            meas(m)%pressure    = to_field ( cur%meas%pressure )
            meas(m)%height      = to_field ( cur%meas%height )
            meas(m)%temperature = to_field ( cur%meas%temperature )
            meas(m)%dew_point   = to_field ( cur%meas%dew_point )
            meas(m)%speed       = to_field ( cur%meas%speed )
            meas(m)%direction   = to_field ( cur%meas%direction )
            meas(m)%u           = to_field ( cur%meas%u )
            meas(m)%v           = to_field ( cur%meas%v )
            meas(m)%rh          = to_field ( cur%meas%rh )
            meas(m)%thickness   = to_field ( cur%meas%thickness )
            reports(i)%n_meas = reports(i)%n_meas + 1
            cur => cur%next
         END DO
      END DO
   END SUBROUTINE og_flatten_reports

   !  What the QC and OA of a time period change in the list: flags (every field of every
   !  measurement), the dew point the DEWPOINT slabs derive, ground%slp%qc, info%discard.
   SUBROUTINE og_unflatten_reports ( obs , number_of_obs , reports , meas )
      TYPE ( report ) , INTENT ( INOUT ) , DIMENSION ( : ) :: obs
      INTEGER , INTENT ( IN ) :: number_of_obs
      TYPE ( og_report ) , INTENT ( IN ) , DIMENSION ( : ) :: reports
      TYPE ( og_meas ) , INTENT ( IN ) , DIMENSION ( : ) :: meas
      TYPE ( measurement ) , POINTER :: cur
      INTEGER :: i , m

      DO i = 1 , number_of_obs
         obs(i)%info%discard  = reports(i)%discard .NE. 0
         obs(i)%ground%slp%qc = reports(i)%slp%qc
         m = reports(i)%first_meas
         cur => obs(i)%surface
         DO WHILE ( ASSOCIATED ( cur ) )
            m = m + 1
            cur%meas%pressure%qc    = meas(m)%pressure%qc
            cur%meas%height%qc      = meas(m)%height%qc
            cur%meas%temperature%qc = meas(m)%temperature%qc
            cur%meas%dew_point%data = meas(m)%dew_point%data
            cur%meas%dew_point%qc   = meas(m)%dew_point%qc
            cur%meas%speed%qc       = meas(m)%speed%qc
            cur%meas%direction%qc   = meas(m)%direction%qc
            cur%meas%u%qc           = meas(m)%u%qc
            cur%meas%v%qc           = meas(m)%v%qc
            cur%meas%rh%qc          = meas(m)%rh%qc
            cur => cur%next
         END DO
      END DO
   END SUBROUTINE og_unflatten_reports

END MODULE obsgrid_gpu_reports
