# Builds libzseek_b200/libzseek_b200.so: C host reader + C-ABI launch layer + sm_100a kernels.
NVCC     ?= nvcc
CC       ?= gcc
ARCH     := -gencode arch=compute_100a,code=sm_100a
EXTRA    ?=
NVFLAGS  := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-fvisibility=hidden,-Wall $(EXTRA)
CFLAGS   := -std=c11 -O2 -g -fPIC -fvisibility=hidden -Wall -Wextra -Iinclude
SRC      := libzseek_b200/csrc
OUT      := libzseek_b200/libzseek_b200.so

all: $(OUT)

$(SRC)/zsk_cuda.o: $(SRC)/zsk_cuda.cu $(wildcard $(SRC)/*.cuh) $(wildcard $(SRC)/*.h)
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(SRC)/reader.o: $(SRC)/reader.c $(wildcard $(SRC)/*.h) $(wildcard include/*.h)
	$(CC) $(CFLAGS) -c $< -o $@

$(OUT): $(SRC)/zsk_cuda.o $(SRC)/reader.o
	$(NVCC) $(ARCH) -cudart shared -shared -o $@ $^ -lpthread

oracle:
	$(MAKE) -C oracle

# experiment build: make alt EXTRA="-DZSK_LZ4_MIN_CTAS=16"  ->  libzseek_b200/libzseek_b200_alt.so (select with ZSEEK_B200_LIB)
alt:
	$(NVCC) $(NVFLAGS) -c $(SRC)/zsk_cuda.cu -o $(SRC)/zsk_cuda_alt.o
	$(CC) $(CFLAGS) -c $(SRC)/reader.c -o $(SRC)/reader_alt.o
	$(NVCC) $(ARCH) -cudart shared -shared -o libzseek_b200/libzseek_b200_alt.so $(SRC)/zsk_cuda_alt.o $(SRC)/reader_alt.o -lpthread

clean:
	rm -f $(SRC)/*.o $(OUT)

.PHONY: all alt oracle clean
