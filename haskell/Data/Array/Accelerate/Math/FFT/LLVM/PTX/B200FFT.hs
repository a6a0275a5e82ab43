{-# LANGUAGE ForeignFunctionInterface #-}
-- |
-- Module      : Data.Array.Accelerate.Math.FFT.LLVM.PTX.B200FFT
--
-- FFI binding of libb200fft (include/b200fft.h).  It exports the same names the PTX backend of
-- accelerate-fft uses from Hackage @cufft@ (module "Foreign.CUDA.FFT"): 'Handle', 'Type', 'Mode',
-- 'plan1D', 'plan2D', 'plan3D', 'planMany', 'execC2C', 'execZ2Z', 'destroy' -- so that
-- @LLVM/PTX.hs@ and @LLVM/PTX/Plans.hs@ switch libraries by changing ONE import line each
-- (see INTEGRATION.md).  The only intentional difference: there is no 'setStream'; the stream is
-- an argument of 'execC2C' / 'execZ2Z' because plans are immutable (cf. the race noted in
-- SURVEY.md section 5: PTX.hs:121 mutates a cached, shared plan outside the cache lock).
--
-- NOT COMPILED in the build container (no GHC there); kept deliberately small so it can be
-- reviewed by eye against include/b200fft.h.
--
module Data.Array.Accelerate.Math.FFT.LLVM.PTX.B200FFT (

  Handle(..), Type(..), Mode(..), B200FFTException(..),
  plan1D, plan2D, plan3D, planMany,
  execC2C, execZ2Z, execC2CScaled, execZ2ZScaled, execC2CShifted, execZ2ZShifted,
  destroy,

) where

import Control.Exception
import Control.Monad                                                ( when )
import Data.Hashable
import Data.Int
import Foreign.C.String
import Foreign.C.Types
import Foreign.CUDA.Driver.Stream                                   ( Stream(..) )
import Foreign.CUDA.Ptr                                             ( DevicePtr(..) )
import Foreign.Marshal.Alloc
import Foreign.Ptr
import Foreign.Storable

-- | b200fftHandle: an opaque, immutable plan.  (replaces FFT.Handle, PTX/Plans.hs:41,42,50,66)
newtype Handle = Handle { useHandle :: Ptr () }
  deriving (Eq, Show)

-- | Element type of the transform; 'fromEnum' equals cuFFT's values, so the hash salts computed in
-- PTX.hs:142,149,156,163,170 are unchanged.
data Type = C2C | Z2Z
  deriving (Eq, Show)

instance Enum Type where
  fromEnum C2C = 0x29
  fromEnum Z2Z = 0x69
  toEnum 0x29  = C2C
  toEnum 0x69  = Z2Z
  toEnum x     = error ("B200FFT.Type.toEnum: " ++ show x)

-- | The plan caches key on the full (context, shape, type) tuple (haskell/patch/0001, PTX/Plans.hs:41,72).
instance Hashable Type where
  hashWithSalt s = hashWithSalt s . fromEnum

-- | Direction; un-normalised both ways (FFT.hs applies the 'Inverse' scale afterwards).
data Mode = Forward | Inverse
  deriving (Eq, Show)

instance Enum Mode where
  fromEnum Forward = -1
  fromEnum Inverse = 1
  toEnum (-1)      = Forward
  toEnum 1         = Inverse
  toEnum x         = error ("B200FFT.Mode.toEnum: " ++ show x)

data B200FFTException = B200FFTException !Int !String
  deriving Show
instance Exception B200FFTException

-- plan creation may allocate and upload twiddle tables: safe calls
foreign import ccall safe   "b200fftPlan1d"     c_plan1d     :: Ptr (Ptr ()) -> Int64 -> CInt -> Int64 -> IO CInt
foreign import ccall safe   "b200fftPlan2d"     c_plan2d     :: Ptr (Ptr ()) -> Int64 -> Int64 -> CInt -> IO CInt
foreign import ccall safe   "b200fftPlan3d"     c_plan3d     :: Ptr (Ptr ()) -> Int64 -> Int64 -> Int64 -> CInt -> IO CInt
foreign import ccall safe   "b200fftPlanMany1d" c_planMany1d :: Ptr (Ptr ()) -> Int64 -> Int64 -> CInt -> IO CInt
-- exec only enqueues kernels, but it takes the scratch-pool mutex, may create the pool on first use and its first launch
-- loads the CUDA module lazily: a `safe` call keeps the GHC capability and the stop-the-world GC free meanwhile, at a cost
-- that is negligible next to a kernel launch
foreign import ccall safe   "b200fftExec"       c_exec       :: Ptr () -> Ptr () -> Ptr () -> CInt -> Ptr () -> IO CInt
foreign import ccall safe   "b200fftExecScaled" c_execScaled :: Ptr () -> Ptr () -> Ptr () -> CInt -> CDouble -> Ptr () -> IO CInt
foreign import ccall safe   "b200fftExecShifted" c_execShifted :: Ptr () -> Ptr () -> Ptr () -> CInt -> CDouble -> Ptr () -> IO CInt
foreign import ccall safe   "b200fftDestroy"    c_destroy    :: Ptr () -> IO CInt
foreign import ccall unsafe "b200fftErrorString" c_errorString :: CInt -> IO CString

check :: CInt -> IO ()
check 0 = return ()
check s = do
  msg <- peekCString =<< c_errorString s
  throwIO (B200FFTException (fromIntegral s) msg)

mkPlan :: (Ptr (Ptr ()) -> IO CInt) -> IO Handle
mkPlan f = alloca $ \p -> do
  check =<< f p
  Handle <$> peek p

ty :: Type -> CInt
ty = fromIntegral . fromEnum

-- | FFT.plan1D n t batch                                  (PTX.hs:141)
plan1D :: Int -> Type -> Int -> IO Handle
plan1D n t batch = mkPlan $ \p -> c_plan1d p (fromIntegral n) (ty t) (fromIntegral batch)

-- | FFT.plan2D h w t                                      (PTX.hs:148)
plan2D :: Int -> Int -> Type -> IO Handle
plan2D h w t = mkPlan $ \p -> c_plan2d p (fromIntegral h) (fromIntegral w) (ty t)

-- | FFT.plan3D d h w t                                    (PTX.hs:155)
plan3D :: Int -> Int -> Int -> Type -> IO Handle
plan3D d h w t = mkPlan $ \p -> c_plan3d p (fromIntegral d) (fromIntegral h) (fromIntegral w) (ty t)

-- | FFT.planMany [n] Nothing Nothing t batch              (PTX.hs:162,169)
-- Only the shape the reference uses is supported: rank 1, contiguous, default embedding.
planMany :: [Int] -> Maybe ([Int], Int, Int) -> Maybe ([Int], Int, Int) -> Type -> Int -> IO Handle
planMany [n] Nothing Nothing t batch = mkPlan $ \p -> c_planMany1d p (fromIntegral n) (fromIntegral batch) (ty t)
planMany _   _       _       _ _     = throwIO (B200FFTException 16 "planMany: only rank-1 contiguous batches (as used by accelerate-fft)")

exec :: Handle -> Mode -> Stream -> DevicePtr a -> DevicePtr a -> IO ()
exec (Handle h) dir (Stream st) (DevicePtr i) (DevicePtr o) =
  check =<< c_exec h (castPtr i) (castPtr o) (fromIntegral (fromEnum dir)) (castPtr st)

-- | FFT.setStream p s >> FFT.execC2C p dir in out         (PTX.hs:121,123)
execC2C :: Handle -> Mode -> Stream -> DevicePtr a -> DevicePtr a -> IO ()
execC2C = exec

-- | FFT.setStream p s >> FFT.execZ2Z p dir in out         (PTX.hs:121,124)
execZ2Z :: Handle -> Mode -> Stream -> DevicePtr a -> DevicePtr a -> IO ()
execZ2Z = exec

execScaled :: Handle -> Mode -> Double -> Stream -> DevicePtr a -> DevicePtr a -> IO ()
execScaled (Handle h) dir s (Stream st) (DevicePtr i) (DevicePtr o) =
  check =<< c_execScaled h (castPtr i) (castPtr o) (fromIntegral (fromEnum dir)) (realToFrac s) (castPtr st)

-- | Fused-Inverse entries (SURVEY.md 8f-2): multiply by the given factor in the last butterfly pass.
execC2CScaled, execZ2ZScaled :: Handle -> Mode -> Double -> Stream -> DevicePtr a -> DevicePtr a -> IO ()
execC2CScaled = execScaled
execZ2ZScaled = execScaled

-- | The transform followed by DFT/Centre.hs's shift1D/2D/3D (zero frequency in the middle), the half rotation folded
-- into the stores of the last butterfly pass of every axis (SURVEY.md 8f-4); power-of-two extents only -- other
-- extents raise B200FFTException NOT_SUPPORTED and the caller composes `shiftND . fftND` as today.
execC2CShifted, execZ2ZShifted :: Handle -> Mode -> Double -> Stream -> DevicePtr a -> DevicePtr a -> IO ()
execC2CShifted (Handle h) dir s (Stream st) (DevicePtr i) (DevicePtr o) =
  check =<< c_execShifted h (castPtr i) (castPtr o) (fromIntegral (fromEnum dir)) (realToFrac s) (castPtr st)
execZ2ZShifted = execC2CShifted

-- | FFT.destroy                                           (PTX/Plans.hs:80)
-- Safe from a GC finaliser thread; a status other than success is ignored there.
destroy :: Handle -> IO ()
destroy (Handle h) = do
  s <- c_destroy h
  when (s /= 0 && s /= 1) $ check s
