! Replacement for Code/RandomNumbersForMC.f95: same module name (RandomNumbers), same public list
! (RandomNumbersForMC.f95:108-110), same generator (MT19937: Matsumoto & Nishimura, ACM TOMACS 8, 1998, in the
! init_genrand / init_by_array / genrand_int32 / genrand_real1 form of their 2002 reference code), written from the
! published algorithm in 64-bit integer arithmetic.
!
! The ONE functional change: type(randomNumberSequence) remembers the seed it was created from (seed, nSeed).  The CUDA
! integrator does not advance the Mersenne Twister state: it keys one counter-based Philox4x32-10 stream per photon with
! that seed vector -- (/ iseed, batch /) in monteCarloDriver.f95:277, (/ batch, iseed /) in planeParallel.f95:207 -- and
! the twister's state after seeding holds no recoverable trace of it.  Everything else that draws from a sequence on
! the host (getRandomReal and friends) behaves exactly like the reference's module.
!
! NOT COMPILED in the build environment (no Fortran compiler exists there); the same generator is restated in C in
! oracle/i3rc_oracle.c (orc_mt_*), which the test-suite pins against the published known-answer vector and numpy.
module RandomNumbers
  implicit none
  private

  integer, parameter :: i8 = selected_int_kind(18)
  integer, parameter :: blockSize = 624, M = 397
  integer(i8), parameter :: mask32 = 4294967295_i8, upperBit = 2147483648_i8, lowerBits = 2147483647_i8, &
                            matrixA = 2567483615_i8             ! 0x9908b0df

  ! Public components, like the reference's type (RandomNumbersForMC.f95:99-102), plus the seed that made the sequence
  type randomNumberSequence
    integer                           :: currentElement = blockSize
    integer, dimension(0:blockSize-1) :: state = 0
    integer, dimension(2)             :: seed  = (/ 0, 0 /)   ! what new_RandomNumberSequence was given (first two)
    integer                           :: nSeed = 0            ! 1: scalar seed, 2 or more: vector seed
  end type randomNumberSequence

  interface new_RandomNumberSequence
    module procedure initialize_scalar, initialize_vector
  end interface new_RandomNumberSequence

  public :: randomNumberSequence
  public :: new_RandomNumberSequence, finalize_RandomNumberSequence, &
            getRandomInt, getRandomPositiveInt, getRandomReal, getRandomDouble
contains
  ! ---- 32-bit words live in 64-bit integers while they are worked on; the state array keeps the reference's layout
  !      (default integers holding the same 32 bits, i.e. negative when bit 31 is set)
  elemental function toWord(i) result(w)      ! default integer -> its 32 bits as a non-negative 64-bit integer
    integer, intent(in) :: i
    integer(i8) :: w
    w = iand(int(i, i8), mask32)
  end function toWord
  elemental function fromWord(w) result(i)    ! and back (two's complement)
    integer(i8), intent(in) :: w
    integer :: i
    if (w >= upperBit) then
      i = int(w - 4294967296_i8)
    else
      i = int(w)
    end if
  end function fromWord

  ! init_genrand: state(i) = 1812433253 * (state(i-1) xor (state(i-1) >> 30)) + i   (mod 2^32)
  subroutine seedState(state, s)
    integer, dimension(0:blockSize-1), intent(out) :: state
    integer(i8), intent(in) :: s
    integer(i8) :: prev
    integer :: i
    prev = iand(s, mask32)
    state(0) = fromWord(prev)
    do i = 1, blockSize - 1
      prev = iand(1812433253_i8 * ieor(prev, ishft(prev, -30)) + int(i, i8), mask32)
      state(i) = fromWord(prev)
    end do
  end subroutine seedState

  function initialize_scalar(seed) result(twister)      ! RandomNumbersForMC.f95:169-185
    integer, intent(in) :: seed
    type(randomNumberSequence) :: twister
    call seedState(twister%state, toWord(seed))
    twister%currentElement = blockSize
    twister%seed  = (/ seed, 0 /)
    twister%nSeed = 1
  end function initialize_scalar

  function initialize_vector(seed) result(twister)      ! init_by_array, RandomNumbersForMC.f95:187-239
    integer, dimension(0:), intent(in) :: seed
    type(randomNumberSequence) :: twister
    integer :: i, j, k, nKey
    integer(i8) :: cur, prev
    nKey = size(seed)
    call seedState(twister%state, 19650218_i8)
    i = 1; j = 0
    do k = max(blockSize, nKey), 1, -1
      prev = toWord(twister%state(i-1)); cur = toWord(twister%state(i))
      cur = iand(ieor(cur, iand(ieor(prev, ishft(prev, -30)) * 1664525_i8, mask32)) + toWord(seed(j)) + int(j, i8), mask32)
      twister%state(i) = fromWord(cur)
      i = i + 1; j = j + 1
      if (i >= blockSize) then
        twister%state(0) = twister%state(blockSize-1); i = 1
      end if
      if (j >= nKey) j = 0
    end do
    do k = blockSize - 1, 1, -1
      prev = toWord(twister%state(i-1)); cur = toWord(twister%state(i))
      cur = iand(ieor(cur, iand(ieor(prev, ishft(prev, -30)) * 1566083941_i8, mask32)) - int(i, i8) + 4294967296_i8, mask32)
      twister%state(i) = fromWord(cur)
      i = i + 1
      if (i >= blockSize) then
        twister%state(0) = twister%state(blockSize-1); i = 1
      end if
    end do
    twister%state(0) = fromWord(upperBit)       ! the most significant bit: a non-zero initial state is assured
    twister%currentElement = blockSize
    twister%seed = 0
    twister%seed(1:min(2, nKey)) = seed(0:min(2, nKey)-1)
    twister%nSeed = nKey
  end function initialize_vector

  ! the next block of 624 words
  subroutine nextState(twister)
    type(randomNumberSequence), intent(inout) :: twister
    integer :: k
    integer(i8) :: y
    do k = 0, blockSize - 1
      y = ior(iand(toWord(twister%state(k)), upperBit), iand(toWord(twister%state(mod(k+1, blockSize))), lowerBits))
      y = ieor(ishft(y, -1), toWord(twister%state(mod(k+M, blockSize))))
      if (iand(toWord(twister%state(mod(k+1, blockSize))), 1_i8) /= 0) y = ieor(y, matrixA)
      twister%state(k) = fromWord(y)
    end do
    twister%currentElement = 0
  end subroutine nextState
  ! (the twist of word k uses the OLD words k+1 and the old or new word k+M exactly as the in-place C loop does: words
  !  k+1 .. 623 are still old when word k is written, and word k+M-624 < k is already new)

  function getRandomInt(twister)                         ! genrand_int32, RandomNumbersForMC.f95:243-257
    type(randomNumberSequence), intent(inout) :: twister
    integer :: getRandomInt
    integer(i8) :: y
    if (twister%currentElement >= blockSize) call nextState(twister)
    y = toWord(twister%state(twister%currentElement))
    twister%currentElement = twister%currentElement + 1
    y = ieor(y, ishft(y, -11))
    y = ieor(y, iand(ishft(y, 7),  2636928640_i8))       ! 0x9d2c5680
    y = ieor(y, iand(ishft(y, 15), 4022730752_i8))       ! 0xefc60000
    y = ieor(y, ishft(y, -18))
    getRandomInt = fromWord(iand(y, mask32))
  end function getRandomInt

  function getRandomPositiveInt(twister)                 ! genrand_int31
    type(randomNumberSequence), intent(inout) :: twister
    integer :: getRandomPositiveInt
    getRandomPositiveInt = int(ishft(toWord(getRandomInt(twister)), -1))
  end function getRandomPositiveInt

  function getRandomDouble(twister)                      ! genrand_real1: [0,1] with 32-bit resolution
    type(randomNumberSequence), intent(inout) :: twister
    double precision :: getRandomDouble
    getRandomDouble = dble(toWord(getRandomInt(twister))) / 4294967295.0d0
  end function getRandomDouble

  function getRandomReal(twister)                        ! RandomNumbersForMC.f95:292-299
    type(randomNumberSequence), intent(inout) :: twister
    real :: getRandomReal
    getRandomReal = real(getRandomDouble(twister))
  end function getRandomReal

  subroutine finalize_RandomNumberSequence(twister)
    type(randomNumberSequence), intent(inout) :: twister
    twister%currentElement = blockSize
    twister%state(:) = 0
    twister%seed(:)  = 0
    twister%nSeed    = 0
  end subroutine finalize_RandomNumberSequence
end module RandomNumbers
